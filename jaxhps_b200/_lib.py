"""ctypes binding of ``libhps_b200.so`` (C ABI in ``include/hps_b200.h``).

PyTorch is used only as the device-memory / stream provider: tensors are allocated with
``torch.empty(..., device="cuda")`` and handed to the library as raw pointers together with
the current CUDA stream.  There is NO CPU fallback: if the shared library is missing or no
CUDA device is present, every stage function raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import numpy as np
import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhps_b200.so")
_lib: Optional[ctypes.CDLL] = None

_i = ctypes.c_int
_l = ctypes.c_int64
_d = ctypes.c_double
_p = ctypes.c_void_p
_sz = ctypes.c_size_t

_SIGNATURES = {
    "hps_version": (_i, []),
    "hps_last_error_string": (ctypes.c_char_p, []),
    "hps_prof_enable": (_i, [_i]),
    "hps_prof_read": (_i, [_p, ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(_l), ctypes.POINTER(_l)]),
    "hps_prof_timeline": (_i, [ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(_i), ctypes.POINTER(_i),
                          ctypes.POINTER(_i), _l, ctypes.POINTER(_l)]),
    "hps_dgemm_strided_batched": (_i, [_p, _i, _i, _i, _d, _p, _l, _l, _p, _l, _l, _d, _p, _l, _l, _i]),
    "hps_lu_solve_workspace": (_i, [_i, _i, ctypes.POINTER(_sz)]),
    "hps_lu_solve": (_i, [_p, _i, _i, _p, _l, _l, _i, ctypes.POINTER(_p), ctypes.POINTER(_l),
                          ctypes.POINTER(_l), ctypes.POINTER(_i), _p, _sz, _p]),
    "hps_local_solve_dtn_workspace": (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_local_solve_dtn": (_i, [_p, _i, _i, _i, _i, _i, ctypes.c_char_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hps_merge_oct_dtn_level_workspace": (_i, [_i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_merge_oct_dtn_level": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "hps_merge_oct_dtn_root_cols": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "hps_root_pack_oct": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "hps_root_solve_oct_workspace": (_i, [_i, ctypes.POINTER(_sz)]),
    "hps_root_solve_oct": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hps_root_assemble_oct": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "hps_root_assemble_panels": (_i, [_p, _i, _i, _i, ctypes.POINTER(_i), _p, _p, _p, _p, _p, _p]),
    "hps_root_solve_panels": (_i, [_p, _i, _i, _i, ctypes.POINTER(_i), _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hps_root_panels_structure": (_i, [_i, ctypes.POINTER(_i), _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "hps_lu_dist_buffer_doubles": (_i, [_i, ctypes.POINTER(_sz)]),
    "hps_lu_dist_factor_pack": (_i, [_p, _i, _p, _l, _i, _p, _sz, _p, _p]),
    "hps_lu_dist_unpack": (_i, [_p, _i, _p, _l, _i, _p, _sz, _p]),
    "hps_lu_dist_update": (_i, [_p, _i, _p, _l, _i, _i, _i, _i, _i, _p, _sz]),
    "hps_lu_dist_solve": (_i, [_p, _i, _p, _l, _i, ctypes.POINTER(_p), ctypes.POINTER(_l), ctypes.POINTER(_i), _p, _sz]),
    "hps_memcpy_d2d": (_i, [_p, _p, _p, _sz]),
    "hps_comm_create": (_i, [_i, _i, ctypes.POINTER(_p)]),
    "hps_comm_destroy": (_i, [_p]),
    "hps_comm_reserve": (_i, [_p, _sz, ctypes.POINTER(_i)]),
    "hps_comm_detach": (_i, [_p]),
    "hps_comm_export": (_i, [_p, _p]),
    "hps_comm_attach": (_i, [_p, _p]),
    "hps_lu_dist_segment_bytes": (_i, [_i, ctypes.POINTER(_sz)]),
    "hps_lu_dist_matrix_ptr": (_i, [_p, _i, ctypes.POINTER(_p)]),
    "hps_lu_dist_run": (_i, [_p, _p, _i, _i, ctypes.POINTER(_p), ctypes.POINTER(_l), ctypes.POINTER(_i), _p, _sz, _p]),
    "hps_lu_dist_run_structured": (_i, [_p, _p, _i, _i, ctypes.POINTER(_p), ctypes.POINTER(_l), ctypes.POINTER(_i), _i, _i,
                                        ctypes.POINTER(_i), _p, _sz, _p]),
    "hps_lu_set_speculative": (_i, [_i]),
    "hps_refine_check_workspace": (_i, [_i, _i, ctypes.POINTER(_sz)]),
    "hps_refine_check": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz]),
    "hps_gemv_t_strided_batched": (_i, [_p, _i, _i, _i, _d, _p, _l, _l, _p, _l, _l, _d, _p, _l, _l, _i, _i]),
    "hps_root_cols_structure": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "hps_lu_dist_apply": (_i, [_p, _p, _i, _i, ctypes.POINTER(_p), ctypes.POINTER(_l), ctypes.POINTER(_i), _p, _sz]),
    "hps_down_oct_scatter": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "hps_merge_quad_dtn_level_workspace": (_i, [_i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_merge_quad_dtn_level": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "hps_local_solve_2d_iti_workspace": (_i, [_i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_local_solve_2d_iti": (_i, [_p, _i, _i, _i, _i, ctypes.c_char_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p, _p]),
    "hps_merge_quad_iti_level_workspace": (_i, [_i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_merge_quad_iti_level": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "hps_merge_quad_dtn_level_nosource": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hps_merge_quad_iti_level_nosource": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hps_up_gather_quad": (_i, [_p, _i, _i, _i, _p, _p, _p, _i]),
    "hps_up_gather_quad_iti": (_i, [_p, _i, _i, _i, _p, _p, _p, _i, ctypes.POINTER(_i)]),
    "hps_zgemm_strided_batched": (_i, [_p, _i, _i, _i, _d, _p, _l, _l, _p, _l, _d, _p, _l, _l, _i, _p]),
    "hps_zgesv_workspace": (_i, [_i, _i, ctypes.POINTER(_sz)]),
    "hps_zgesv": (_i, [_p, _i, _i, _p, _l, _p, _l, _p, _p, _sz, _p]),
    "hps_down_quad_iti_level": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "hps_leaf_apply_complex": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "hps_down_oct_level": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "hps_down_quad_level": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "hps_leaf_apply": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "hps_interp_from_hps": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "hps_interp_to_hps_workspace": (_i, [_i, _i, _i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_interp_to_hps": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz]),
    "hps_adaptive_compress_workspace": (_i, [_i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_adaptive_compress": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _sz]),
    "hps_merge_adaptive_workspace": (_i, [_i, _i, _i, ctypes.POINTER(_sz)]),
    "hps_merge_adaptive": (_i, [_p, _i, _i, _i, ctypes.POINTER(_p), ctypes.POINTER(_p), ctypes.POINTER(_i), _i, _p, _i, _p,
                                _p, _p, _p, _p, _i, _i, ctypes.POINTER(_i), _i, _i, _p, _sz, _p]),
    "hps_merge_adaptive_assemble": (_i, [_p, _i, _i, _i, ctypes.POINTER(_p), ctypes.POINTER(_p), ctypes.POINTER(_i), _i, _p, _i,
                                         _p, _p, _p, _p, _i, _i]),
    "hps_down_adaptive": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, ctypes.POINTER(_p), _i, _p, _p, _p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class HpsLibraryError(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Fails loudly; never falls back to a CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise HpsLibraryError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  jaxhps_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hps_last_error_string().decode(errors="replace")
        raise HpsLibraryError(f"{what} failed with code {rc}: {msg}")


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise HpsLibraryError(
            "jaxhps_b200 runs its hot path as CUDA kernels only; no CUDA device is visible "
            "(there is no CPU fallback — use the oracle under oracle/ for CPU checks)."
        )
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise HpsLibraryError(f"compute device must be a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def is_host(host_device) -> bool:
    """``host_device`` follows the reference's meaning: where results are returned.  ``None`` or
    ``"cpu"`` -> NumPy arrays on the host; a CUDA device -> results stay resident as torch tensors."""
    if host_device is None:
        return True
    return torch.device(host_device).type == "cpu"


def to_device(x, dev: torch.device, dtype=torch.float64) -> torch.Tensor:
    """NumPy / torch (any device) -> contiguous tensor on ``dev``; host arrays go through pinned memory.
    A complex array handed to a real-valued (DtN) stage is an error, not a silent drop of the imaginary part
    (the reference would promote to complex; the DtN kernels are FP64-only — use the ItI path for complex fields)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    if t.dtype != dtype:
        if t.is_complex() and not dtype.is_complex:
            raise ValueError("complex-valued input reached a real-valued (DtN) stage; complex coefficient fields, "
                             "sources and boundary data are supported on the ItI path only")
        t = t.to(dtype)
    if t.device != dev:
        if t.device.type == "cpu" and t.numel() > 1 << 16 and not t.is_pinned():
            t = t.contiguous().pin_memory()
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


def to_result(t: torch.Tensor, host_device):
    """Device tensor -> what the caller asked for (NumPy on the host, or the tensor itself)."""
    if is_host(host_device):
        if not isinstance(t, torch.Tensor):
            return np.asarray(t)
        if t.is_cuda and t.numel() * t.element_size() >= PINNED_RESULT_MIN_BYTES:
            # large operators (Y, S_lst: GBs) land in page-locked host memory: the copy runs at PCIe speed instead of
            # the pageable rate, and `to_device` recognises the array as pinned when `solve` brings it back (the blocks
            # are recycled by torch's pinned-memory cache from one build to the next)
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            return buf.numpy()
        return t.cpu().numpy()
    dev = torch.device(host_device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if not isinstance(t, torch.Tensor):
        return to_device(t, dev)
    return t if t.device == dev else t.to(dev)


#: results at least this large are returned to the host through page-locked memory (see `to_result`)
PINNED_RESULT_MIN_BYTES = 8 << 20

#: most matrices / merges / nodes one C-ABI call accepts (CUDA grid-dimension limit)
MAX_BATCH = 65535


class Workspace:
    """Grow-only device scratch buffers handed to the library (which never allocates): one per
    (device, CUDA stream), so that calls issued on different streams or devices — from one host thread or
    several — never share scratch memory."""

    def __init__(self):
        self._bufs = {}
        self._lock = threading.Lock()

    def get(self, nbytes: int, dev: torch.device) -> torch.Tensor:
        dev = torch.device(dev)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (idx, torch.cuda.current_stream(idx).cuda_stream)
        with self._lock:
            buf = self._bufs.get(key)
            if buf is None or buf.numel() < nbytes:
                self._bufs.pop(key, None)
                del buf
                buf = torch.empty(int(nbytes), dtype=torch.uint8, device=torch.device("cuda", idx))
                self._bufs[key] = buf
            return buf

    def release(self):
        with self._lock:
            self._bufs.clear()


WORKSPACE = Workspace()


def workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    """Scratch buffer of at least ``nbytes`` for the current stream of ``dev``."""
    return WORKSPACE.get(nbytes, dev)


class AssumptionViolated(Exception):
    """A structural shortcut of the library did not apply to this matrix (``info < 0``: the factorisation of a merge
    matrix needed row interchanges below a diagonal block).  The operation is repeated with the shortcut off."""


def check_info(info: torch.Tensor, what: str) -> None:
    """LAPACK-style singularity report (one host sync; the stages call it once per level)."""
    bad = torch.nonzero(info)
    if bad.numel():
        pos = torch.nonzero(info > 0)
        if pos.numel():
            k = int(pos[0, 0])
            raise np.linalg.LinAlgError(f"{what}: exact zero pivot in matrix {k} at column {int(info[k])}")
        raise AssumptionViolated(what)


_SPEC_DEFAULT = 0 if os.environ.get("HPS_LU_SPEC", "1") == "0" else 1


class speculation:
    """Context manager: switch the library's speculative (no-pivot) block columns on or off for the calls inside."""

    def __init__(self, on: bool):
        self.on = 1 if on else 0

    def __enter__(self):
        load().hps_lu_set_speculative(self.on)

    def __exit__(self, *exc):
        load().hps_lu_set_speculative(_SPEC_DEFAULT)
        return False


def with_pivoting_fallback(fn):
    """Run ``fn()``; if the library reports that its no-pivoting speculation failed, run it again with the
    speculation switched off (``hps_lu_set_speculative``)."""
    try:
        return fn()
    except AssumptionViolated:
        lib = load()
        lib.hps_lu_set_speculative(0)
        try:
            return fn()
        finally:
            lib.hps_lu_set_speculative(_SPEC_DEFAULT)
