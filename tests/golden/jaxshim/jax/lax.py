import numpy as _np
from ._core import wrap_out as _wrap_out


def cond(pred, true_fn, false_fn, *operands):
    return true_fn(*operands) if bool(pred) else false_fn(*operands)


def switch(index, branches, *operands):
    return branches[int(index)](*operands)


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def dynamic_slice(x, start, sizes):
    sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, sizes))
    return _wrap_out(_np.asarray(x)[sl])
