"""down_pass on the B200: API mirror of `src/jaxhps/down_pass/_uniform_3D_DtN.py:8-113` and
`down_pass/_uniform_2D_DtN.py:7-122`; ``hps_down_*_level`` per tree level then
``hps_leaf_apply``."""
from __future__ import annotations

import torch

from . import _lib


def down_levels(g_cur: torch.Tensor, S_dev, g_dev, dim: int, dev) -> torch.Tensor:
    """Push boundary data ``g_cur (n_nodes, n_ext, n_src)`` through the given levels (root-most
    last in the lists); returns ``(n_nodes * n_child^levels, n_face*m, n_src)``."""
    lib = _lib.load()
    n_child = 8 if dim == 3 else 4
    n_face = 6 if dim == 3 else 4
    n_slot = 12 if dim == 3 else 4
    down_fn = lib.hps_down_oct_level if dim == 3 else lib.hps_down_quad_level
    n_src = g_cur.shape[-1]
    for level in range(len(S_dev) - 1, -1, -1):
        S = S_dev[level]
        gt = g_dev[level].reshape(S.shape[0], S.shape[1], n_src)
        n_nodes, n_int, n_ext = S.shape
        m = n_int // n_slot
        if g_cur.shape[0] != n_nodes or g_cur.shape[1] != n_ext:
            raise ValueError(
                f"level {level}: boundary data of shape {tuple(g_cur.shape)} does not match S {tuple(S.shape)}"
            )
        out = torch.empty((n_nodes * n_child, n_face * m, n_src), dtype=torch.float64, device=dev)
        ws = torch.empty((n_nodes, n_int, n_src), dtype=torch.float64, device=dev)
        step = min(n_nodes, _lib.MAX_BATCH)  # the C ABI takes at most MAX_BATCH nodes per call
        for s0 in range(0, n_nodes, step):
            s1 = min(n_nodes, s0 + step)
            rc = down_fn(_lib.stream_ptr(), s1 - s0, m, n_src, _lib.ptr(S[s0:s1]), _lib.ptr(g_cur[s0:s1]),
                         _lib.ptr(gt[s0:s1]), _lib.ptr(out[n_child * s0:n_child * s1]), _lib.ptr(ws[s0:s1]))
            _lib.check(rc, "hps_down_level")
        g_cur = out
    return g_cur


def leaf_apply(Y: torch.Tensor, g_leaf: torch.Tensor, v: torch.Tensor, dev) -> torch.Tensor:
    """``u = Y g + v`` on every leaf; all arguments 3-D device tensors."""
    lib = _lib.load()
    n_leaves, n_c, n_g = Y.shape
    n_src = g_leaf.shape[-1]
    u = torch.empty((n_leaves, n_c, n_src), dtype=torch.float64, device=dev)
    rc = lib.hps_leaf_apply(_lib.stream_ptr(), n_leaves, n_c, n_g, n_src, _lib.ptr(Y), _lib.ptr(g_leaf),
                            _lib.ptr(v.reshape(n_leaves, n_c, n_src)), _lib.ptr(u))
    _lib.check(rc, "hps_leaf_apply")
    return u


def _down_pass(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr, dim: int, device, host_device):
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    n_child = 8 if dim == 3 else 4
    n_face = 6 if dim == 3 else 4
    n_slot = 12 if dim == 3 else 4
    down_fn = lib.hps_down_oct_level if dim == 3 else lib.hps_down_quad_level
    with torch.cuda.device(dev):
        S_dev = [_lib.to_device(S, dev) for S in S_lst]
        g_dev = [_lib.to_device(g, dev) for g in g_tilde_lst]
        # the reference's multi-source detection (3D: `_uniform_3D_DtN.py:70`; 2D: `_uniform_2D_DtN.py:67`)
        if dim == 3:
            multi = len(g_dev) > 1 and g_dev[0].ndim == 3
            # root entries are stored without a batch axis in 3D
            S_dev[-1] = S_dev[-1].unsqueeze(0)
            g_dev[-1] = g_dev[-1].unsqueeze(0)
        else:
            multi = g_dev[0].ndim == 3
        bd = _lib.to_device(boundary_data, dev)
        if multi and bd.ndim == 1:
            raise ValueError("For multi-source downward pass, need to specify boundary data for each source.")
        n_src = bd.shape[-1] if multi else 1
        g_cur = bd.reshape(1, -1, n_src).contiguous()
        g_cur = down_levels(g_cur, S_dev, g_dev, dim, dev)
        if Y_arr is None:
            res = g_cur if multi else g_cur[..., 0]
            return _lib.to_result(res, host_device)
        Y = _lib.to_device(Y_arr, dev)
        v = _lib.to_device(v_arr, dev)
        u = leaf_apply(Y, g_cur, v, dev)
        return _lib.to_result(u if multi else u[..., 0], host_device)


def down_pass_uniform_3D_DtN(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr, device=None, host_device=None):
    """Propagate Dirichlet data from the root to the leaves and evaluate ``u = Y g + v``
    (reference `down_pass/_uniform_3D_DtN.py:8-113`).  Returns ``(n_leaves, p^3[, n_src])``."""
    return _down_pass(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr, 3, device, host_device)


def down_pass_uniform_2D_DtN(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr, device=None, host_device=None):
    """2D analogue (reference `down_pass/_uniform_2D_DtN.py:7-122`); ``Y_arr=None`` returns the
    leaves' boundary data instead of the solution (`:111-112`)."""
    return _down_pass(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr, 2, device, host_device)


def down_pass_uniform_2D_ItI(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr, device=None, host_device=None):
    """ItI down pass, complex128 (reference `down_pass/_uniform_2D_ItI.py:8-118`): incoming impedance
    data is pushed from the root to the leaves, then ``u = Y g + v``.  ``Y_arr=None`` returns the
    leaves' incoming impedance data (`:105-106`)."""
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        S_dev = [_lib.to_device(S, dev, dtype=torch.complex128) for S in S_lst]
        g_dev = [_lib.to_device(g, dev, dtype=torch.complex128) for g in g_tilde_lst]
        multi = bool(len(g_dev)) and g_dev[0].ndim == 3
        bd = _lib.to_device(boundary_data, dev, dtype=torch.complex128)
        if multi and bd.ndim == 1:
            raise ValueError("For multi-source downward pass, need to specify boundary data for each source.")
        n_src = bd.shape[-1] if multi else 1
        g_cur = bd.reshape(1, -1, n_src).contiguous()
        c128 = dict(dtype=torch.complex128, device=dev)
        for level in range(len(S_dev) - 1, -1, -1):
            S = S_dev[level]
            n_nodes, n, _ = S.shape
            m = n // 8
            gt = g_dev[level].reshape(n_nodes, n, n_src)
            if g_cur.shape[0] != n_nodes or g_cur.shape[1] != n:
                raise ValueError(f"level {level}: boundary data {tuple(g_cur.shape)} does not match S {tuple(S.shape)}")
            out = torch.empty((n_nodes * 4, 4 * m, n_src), **c128)
            ws = torch.empty(n_nodes * 48 * m * n_src, dtype=torch.float64, device=dev)
            rc = lib.hps_down_quad_iti_level(_lib.stream_ptr(), n_nodes, m, n_src, _lib.ptr(S), _lib.ptr(g_cur), _lib.ptr(gt),
                                             _lib.ptr(out), _lib.ptr(ws))
            _lib.check(rc, "hps_down_quad_iti_level")
            g_cur = out
        if Y_arr is None:
            return _lib.to_result(g_cur if multi else g_cur[..., 0], host_device)
        Y = _lib.to_device(Y_arr, dev, dtype=torch.complex128)
        v = _lib.to_device(v_arr, dev, dtype=torch.complex128)
        n_leaves, n_c, n_g = Y.shape
        u = torch.empty((n_leaves, n_c, n_src), **c128)
        ws = torch.empty(n_leaves * 4 * n_g * n_src, dtype=torch.float64, device=dev)
        rc = lib.hps_leaf_apply_complex(_lib.stream_ptr(), n_leaves, n_c, n_g, n_src, _lib.ptr(Y), _lib.ptr(g_cur),
                                        _lib.ptr(v.reshape(n_leaves, n_c, n_src)), _lib.ptr(u), _lib.ptr(ws))
        _lib.check(rc, "hps_leaf_apply_complex")
        return _lib.to_result(u if multi else u[..., 0], host_device)



def __getattr__(name):  # the adaptive-tree stages live in adaptive.py (imported lazily: it imports this module)
    if name in ('down_pass_adaptive_2D_DtN', 'down_pass_adaptive_3D_DtN'):
        from . import adaptive

        return getattr(adaptive, name)
    raise AttributeError(name)
