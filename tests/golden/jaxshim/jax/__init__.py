"""`jax` shim — see _core.py.  TEST INFRASTRUCTURE ONLY."""
import numpy as _np
from ._core import ShimArray as Array, Device, DEVICES, wrap_out as _wrap_out
from . import numpy, lax, tree_util, typing, sharding, debug, tree, core  # noqa: F401

Dtype = _np.dtype


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def devices(kind=None):
    return list(DEVICES)


def device_put(x, device=None):
    if x is None:
        return None
    if isinstance(x, (list, tuple)):
        return type(x)(device_put(y, device) for y in x)
    return _wrap_out(_np.asarray(x))


def jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def clear_caches():
    return None


def _moveaxis_take(x, axis, i):
    return _np.take(x, i, axis=axis)


def vmap(fn, in_axes=0, out_axes=0):
    def batched(*args):
        if isinstance(in_axes, int) or in_axes is None:
            axes = [in_axes] * len(args)
        else:
            axes = list(in_axes)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            sl = [
                a if ax is None else _wrap_out(_np.take(_np.asarray(a), i, axis=ax))
                for a, ax in zip(args, axes)
            ]
            outs.append(fn(*sl))
        if isinstance(outs[0], (tuple, list)):
            k = len(outs[0])
            oaxes = [out_axes] * k if isinstance(out_axes, int) else list(out_axes)
            return tuple(
                _wrap_out(_np.stack([_np.asarray(o[j]) for o in outs], axis=oaxes[j]))
                for j in range(k)
            )
        oax = out_axes if isinstance(out_axes, int) else out_axes[0]
        return _wrap_out(_np.stack([_np.asarray(o) for o in outs], axis=oax))

    return batched
