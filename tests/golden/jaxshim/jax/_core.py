"""NumPy-backed stand-in for the tiny slice of the JAX API that meliao/jaxhps uses.

TEST INFRASTRUCTURE ONLY.  JAX is not installable in the build image (no network),
so the unmodified reference under /root/reference cannot be imported as is.  This
shim lets `tests/golden/make_golden.py` import and *execute the reference's own
Python code* (its index maps, its merge assembly, its stage loops) on NumPy so the
outputs can be frozen as golden vectors and the oracle can be pinned to them.
Arithmetic therefore runs through NumPy/LAPACK instead of jaxlib/XLA; everything
else (operation order, index bookkeeping) is the reference's.  Nothing in the
product package imports this.
"""
import numpy as np


class _AtIndexer:
    __slots__ = ("_arr", "_idx")

    def __init__(self, arr, idx):
        self._arr = arr
        self._idx = idx

    def _copy(self):
        return np.array(self._arr, copy=True).view(ShimArray)

    def set(self, value):
        out = self._copy()
        out[self._idx] = value
        return out

    def add(self, value):
        out = self._copy()
        np.add.at(out, self._idx, value)
        return out

    def mul(self, value):
        out = self._copy()
        np.multiply.at(out, self._idx, value)
        return out


class _At:
    __slots__ = ("_arr",)

    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIndexer(self._arr, idx)


class ShimArray(np.ndarray):
    """ndarray with the handful of jax.Array methods the reference calls."""

    def __array_finalize__(self, obj):
        pass

    @property
    def at(self):
        return _At(self)

    def delete(self):
        return None

    def block_until_ready(self):
        return self

    def devices(self):
        return {DEVICES[0]}

    def __getitem__(self, idx):
        # JAX clamps out-of-range scalar indices instead of raising; the reference's
        # lax.cond operands rely on that (local_solve/_uniform_2D_DtN.py:165-171).
        if isinstance(idx, (int, np.integer)) and self.ndim > 0:
            n = self.shape[0]
            if idx >= n:
                idx = n - 1
            elif idx < -n:
                idx = 0
        return super().__getitem__(idx)

    def __iter__(self):
        for i in range(self.shape[0]):
            yield super().__getitem__(i)

    def __hash__(self):  # jit static args are sometimes hashed
        return id(self)


def wrap_out(x):
    if isinstance(x, np.ndarray):
        return x.view(ShimArray)
    if isinstance(x, tuple):
        return tuple(wrap_out(y) for y in x)
    if isinstance(x, list):
        return [wrap_out(y) for y in x]
    return x


def wrap_fn(fn):
    def inner(*a, **k):
        return wrap_out(fn(*a, **k))

    inner.__name__ = getattr(fn, "__name__", "wrapped")
    return inner


class Device:
    def __init__(self, kind="cpu"):
        self.device_kind = kind
        self.platform = kind
        self.id = 0

    def memory_stats(self):
        return {"bytes_limit": 1 << 40}

    def __repr__(self):
        return "ShimCpuDevice(id=0)"


DEVICES = [Device("cpu")]
