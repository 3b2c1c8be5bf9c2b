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


def jit(fn=None, static_argnums=(), static_argnames=(), **kwargs):
    """Identity "compiler"; like the real one it hands plain NumPy inputs to the function as arrays of the
    array type (so that `.at[...]` works on them), leaving static arguments alone."""
    import functools

    static_pos = (static_argnums,) if isinstance(static_argnums, int) else tuple(static_argnums or ())
    static_kw = (static_argnames,) if isinstance(static_argnames, str) else tuple(static_argnames or ())

    def deco(f):
        @functools.wraps(f)
        def wrapped(*args, **kw):
            args = tuple(_wrap_out(a) if (type(a) is _np.ndarray and i not in static_pos) else a for i, a in enumerate(args))
            kw = {k: (_wrap_out(v) if (type(v) is _np.ndarray and k not in static_kw) else v) for k, v in kw.items()}
            return f(*args, **kw)

        return wrapped

    if fn is None:
        return deco
    return deco(fn)


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
