class Mesh:
    def __init__(self, devices, axis_names=()):
        self.devices = devices
        self.axis_names = axis_names
