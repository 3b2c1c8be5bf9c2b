class Tracer:  # nothing is ever traced by the shim
    pass
