"""local_solve_stage on the B200: API mirror of the reference's stage functions
(`src/jaxhps/local_solve/_uniform_3D_DtN.py:13-106`, `_uniform_2D_DtN.py:9-102`) whose device
body is ``hps_local_solve_dtn`` in libhps_b200.so."""
from __future__ import annotations

import ctypes
import logging
from typing import Tuple

import numpy as np
import torch

from . import _lib

_ORDER_3D = ("D_xx", "D_xy", "D_yy", "D_xz", "D_yz", "D_zz", "D_x", "D_y", "D_z", "I")
_ORDER_2D = ("D_xx", "D_xy", "D_yy", "D_x", "D_y", "I")

#: upper bound on the scratch one local-solve call may take (bytes)
MAX_WORKSPACE_BYTES = 24 << 30


def _gather_coeffs(pde_problem, order, dev) -> Tuple[torch.Tensor, bytes]:
    """Stack the non-None coefficient arrays in the reference's fixed order
    (`_uniform_3D_DtN.py:109-145`, `_uniform_2D_DtN.py:105-133`)."""
    arrs = [getattr(pde_problem, f"{name}_coefficients", None) for name in order]
    which = bytes(1 if a is not None else 0 for a in arrs)
    present = [_lib.to_device(a, dev) for a in arrs if a is not None]
    if not present:
        raise ValueError("at least one differential-operator coefficient must be given")
    return torch.stack(present).contiguous(), which


def _gather_coeffs_complex(pde_problem, order, dev):
    """Like :func:`_gather_coeffs` for fields that may be complex (ItI): real parts, imaginary parts
    (``None`` when every field is real) and the presence mask."""
    arrs = [getattr(pde_problem, f"{name}_coefficients", None) for name in order]
    which = bytes(1 if a is not None else 0 for a in arrs)
    present = [a for a in arrs if a is not None]
    if not present:
        raise ValueError("at least one differential-operator coefficient must be given")

    def is_cplx(a):
        return a.is_complex() if isinstance(a, torch.Tensor) else np.iscomplexobj(a)

    def part(a, imag):
        if isinstance(a, torch.Tensor):
            return (a.imag if imag else a.real) if a.is_complex() else (torch.zeros_like(a) if imag else a)
        a = np.asarray(a)
        return np.ascontiguousarray(a.imag if imag else a.real) if np.iscomplexobj(a) else (np.zeros_like(a) if imag else a)

    re = torch.stack([_lib.to_device(part(a, False), dev) for a in present]).contiguous()
    im = None
    if any(is_cplx(a) for a in present):
        im = torch.stack([_lib.to_device(part(a, True), dev) for a in present]).contiguous()
    return re, im, which


def _constants(pde_problem, dev):
    """Device copies of D1, P, Q, cached on the problem object."""
    cache = pde_problem.__dict__.setdefault("_device_constants", {})
    key = str(dev)
    if key not in cache:
        cache[key] = tuple(_lib.to_device(getattr(pde_problem, n), dev) for n in ("D1", "P", "Q"))
    return cache[key]


def _local_solve_dtn(pde_problem, dim: int, device, host_device):
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    dom = pde_problem.domain
    p, q = dom.p, dom.q
    order = _ORDER_3D if dim == 3 else _ORDER_2D
    with torch.cuda.device(dev):
        coeffs, which = _gather_coeffs(pde_problem, order, dev)
        src = _lib.to_device(pde_problem.source, dev)
        multi = src.ndim == 3
        if not multi:
            src = src.unsqueeze(-1)
        n_leaves, n_c, n_src = src.shape
        D1, P, Q = _constants(pde_problem, dev)
        n_g = Q.shape[0]
        Y = torch.empty((n_leaves, n_c, n_g), dtype=torch.float64, device=dev)
        T = torch.empty((n_leaves, n_g, n_g), dtype=torch.float64, device=dev)
        v = torch.empty((n_leaves, n_c, n_src), dtype=torch.float64, device=dev)
        h = torch.empty((n_leaves, n_g, n_src), dtype=torch.float64, device=dev)
        info = torch.zeros(n_leaves, dtype=torch.int32, device=dev)

        # chunk the leaves so that the operator workspace stays bounded
        one = ctypes.c_size_t()
        _lib.check(lib.hps_local_solve_dtn_workspace(dim, 1, p, q, n_src, ctypes.byref(one)), "workspace query")
        free_b, _ = torch.cuda.mem_get_info(dev)
        budget = min(MAX_WORKSPACE_BYTES, int(0.6 * free_b))
        chunk = int(max(1, min(n_leaves, budget // max(1, one.value), 65535)))
        need = ctypes.c_size_t()
        _lib.check(lib.hps_local_solve_dtn_workspace(dim, chunk, p, q, n_src, ctypes.byref(need)), "workspace query")
        ws = _lib.WORKSPACE.get(need.value, dev)
        logging.debug("local_solve: %d leaves in chunks of %d (workspace %.2f GB)", n_leaves, chunk, need.value / 2**30)
        def run():
            for s in range(0, n_leaves, chunk):
                e = min(n_leaves, s + chunk)
                c_chunk = coeffs[:, s:e].contiguous() if (s, e) != (0, n_leaves) else coeffs
                rc = lib.hps_local_solve_dtn(
                    _lib.stream_ptr(), dim, e - s, p, q, n_src, which, _lib.ptr(c_chunk), _lib.ptr(D1), _lib.ptr(P),
                    _lib.ptr(Q), _lib.ptr(src[s:e]), _lib.ptr(Y[s:e]), _lib.ptr(T[s:e]), _lib.ptr(v[s:e]),
                    _lib.ptr(h[s:e]), _lib.ptr(ws), ws.numel(), _lib.ptr(info[s:e]),
                )
                _lib.check(rc, "hps_local_solve_dtn")
            _lib.check_info(info, "local solve")

        _lib.with_pivoting_fallback(run)  # (the stage only reads its inputs: a repeat with full pivoting starts clean)
        if not multi:
            v, h = v[..., 0], h[..., 0]
        return tuple(_lib.to_result(t, host_device) for t in (Y, T, v, h))


def local_solve_stage_uniform_3D_DtN(pde_problem, device=None, host_device=None):
    """Leaf DtN maps for a uniform octree.  Returns ``(Y, T, v, h)`` with shapes
    ``(n, p^3, 6q^2)``, ``(n, 6q^2, 6q^2)``, ``(n, p^3[, n_src])``, ``(n, 6q^2[, n_src])``
    (reference `local_solve/_uniform_3D_DtN.py:13-106`)."""
    return _local_solve_dtn(pde_problem, 3, device, host_device)


def local_solve_stage_uniform_2D_DtN(pde_problem, host_device=None, device=None):
    """Leaf DtN maps for a uniform quadtree; note the reference's argument order
    (`local_solve/_uniform_2D_DtN.py:9-13`)."""
    return _local_solve_dtn(pde_problem, 2, device, host_device)


def local_solve_stage_uniform_2D_ItI(pde_problem, device=None, host_device=None):
    """Leaf impedance-to-impedance maps, complex128.  Returns ``(Y, R, v, h)`` like the reference
    (`local_solve/_uniform_2D_ItI.py:10-117`): ``(n, p^2, 4q)``, ``(n, 4q, 4q)``, ``(n, p^2[, n_src])``,
    ``(n, 4q[, n_src])``."""
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    dom = pde_problem.domain
    p, q = dom.p, dom.q
    with torch.cuda.device(dev):
        coeffs, coeffs_im, which = _gather_coeffs_complex(pde_problem, _ORDER_2D, dev)
        src = _lib.to_device(pde_problem.source, dev, dtype=torch.complex128)
        multi = src.ndim == 3
        if not multi:
            src = src.unsqueeze(-1)
        n_leaves, n_c, n_src = src.shape
        cache = pde_problem.__dict__.setdefault("_device_constants", {})
        key = ("iti", str(dev))
        if key not in cache:
            cache[key] = (_lib.to_device(pde_problem.D1, dev), _lib.to_device(pde_problem.P, dev),
                          _lib.to_device(pde_problem.G, dev, dtype=torch.complex128),
                          _lib.to_device(pde_problem.QH, dev, dtype=torch.complex128))
        D1, P, G, QH = cache[key]
        n_g = QH.shape[0]
        c128 = dict(dtype=torch.complex128, device=dev)
        Y = torch.empty((n_leaves, n_c, n_g), **c128)
        R = torch.empty((n_leaves, n_g, n_g), **c128)
        v = torch.empty((n_leaves, n_c, n_src), **c128)
        h = torch.empty((n_leaves, n_g, n_src), **c128)
        info = torch.zeros(n_leaves, dtype=torch.int32, device=dev)
        one = ctypes.c_size_t()
        _lib.check(lib.hps_local_solve_2d_iti_workspace(1, p, q, n_src, ctypes.byref(one)), "workspace query")
        free_b, _ = torch.cuda.mem_get_info(dev)
        budget = min(MAX_WORKSPACE_BYTES, int(0.6 * free_b))
        chunk = int(max(1, min(n_leaves, budget // max(1, one.value), 65535)))
        need = ctypes.c_size_t()
        _lib.check(lib.hps_local_solve_2d_iti_workspace(chunk, p, q, n_src, ctypes.byref(need)), "workspace query")
        ws = _lib.WORKSPACE.get(need.value, dev)
        for s in range(0, n_leaves, chunk):
            e = min(n_leaves, s + chunk)
            whole = (s, e) == (0, n_leaves)
            c_chunk = coeffs if whole else coeffs[:, s:e].contiguous()
            ci_chunk = None if coeffs_im is None else (coeffs_im if whole else coeffs_im[:, s:e].contiguous())
            rc = lib.hps_local_solve_2d_iti(
                _lib.stream_ptr(), e - s, p, q, n_src, which, _lib.ptr(c_chunk), _lib.ptr(D1), _lib.ptr(P), _lib.ptr(G),
                _lib.ptr(QH), _lib.ptr(src[s:e]), _lib.ptr(Y[s:e]), _lib.ptr(R[s:e]), _lib.ptr(v[s:e]), _lib.ptr(h[s:e]),
                _lib.ptr(ws), ws.numel(), _lib.ptr(info[s:e]), _lib.ptr(ci_chunk),
            )
            _lib.check(rc, "hps_local_solve_2d_iti")
        _lib.check_info(info, "ItI local solve")
        if not multi:
            v, h = v[..., 0], h[..., 0]
        return tuple(_lib.to_result(t, host_device) for t in (Y, R, v, h))



def __getattr__(name):  # same public names as the reference's `jaxhps.local_solve` package; imported lazily (those modules import this one)
    if name in ('local_solve_stage_adaptive_2D_DtN', 'local_solve_stage_adaptive_3D_DtN'):
        from . import adaptive

        return getattr(adaptive, name)
    if name in ('nosource_local_solve_stage_uniform_2D_DtN', 'nosource_local_solve_stage_uniform_2D_ItI'):
        from . import up_pass

        return getattr(up_pass, name)
    raise AttributeError(name)
