"""`jax.numpy` shim: forwards to NumPy, returning ShimArray.  TEST INFRASTRUCTURE ONLY."""
import numpy as _np
from .._core import ShimArray, wrap_fn as _wrap_fn, wrap_out as _wrap_out
from . import linalg, fft  # noqa: F401

ndarray = ShimArray
float64 = _np.float64
float32 = _np.float32
complex128 = _np.complex128
complex64 = _np.complex64
int32 = _np.int32
int64 = _np.int64
bool_ = _np.bool_
pi = _np.pi
nan = _np.nan
inf = _np.inf
newaxis = None
finfo = _np.finfo


def array(x, dtype=None):
    return _wrap_out(_np.array(x, dtype=dtype))


def asarray(x, dtype=None):
    return _wrap_out(_np.asarray(x, dtype=dtype))


def zeros_like(x, dtype=None):
    return _wrap_out(_np.zeros_like(_np.asarray(x), dtype=dtype))


def ones_like(x, dtype=None):
    return _wrap_out(_np.ones_like(_np.asarray(x), dtype=dtype))


def __getattr__(name):
    return _wrap_fn(getattr(_np, name))
