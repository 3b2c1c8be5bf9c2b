"""merge_stage on the B200: API mirror of `src/jaxhps/merge/_uniform_3D_DtN.py:12-124` and
`merge/_uniform_2D_DtN.py:13-203`; one ``hps_merge_*_dtn_level`` call per tree level."""
from __future__ import annotations

import ctypes
import logging

import torch

from . import _lib


def _merge_stage(T_arr, h_arr, l: int, dim: int, device, host_device, return_T: bool, subtree_recomp: bool,
                 return_h: bool = False, n_roots: int = 1):
    """``l`` merge levels over ``n_roots`` independent subtrees stored back to back (``n_roots=1``: the
    reference's whole-tree merge; ``n_roots>1``: the subtrees one GPU owns in the sharded build)."""
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    n_child = 8 if dim == 3 else 4
    n_face = 6 if dim == 3 else 4
    level_fn = lib.hps_merge_oct_dtn_level if dim == 3 else lib.hps_merge_quad_dtn_level
    ws_fn = lib.hps_merge_oct_dtn_level_workspace if dim == 3 else lib.hps_merge_quad_dtn_level_workspace
    with torch.cuda.device(dev):
        T = _lib.to_device(T_arr, dev)
        h = _lib.to_device(h_arr, dev)
        multi = h.ndim == 3
        if not multi:
            h = h.unsqueeze(-1)
        n_src = h.shape[-1]
        if l <= 0 and n_roots == 1 and T.shape[0] == n_child:
            l = 1  # the reference runs l - 1 batched levels plus one final merge, so l = 0 on 4 / 8 operators is l = 1
        if T.shape[0] != n_roots * n_child**l:
            raise ValueError(f"expected {n_roots * n_child**l} leaf operators for l={l}, got {T.shape[0]}")
        S_lst, g_lst = [], []
        for level in range(l, 0, -1):
            n_merges = T.shape[0] // n_child
            m = T.shape[-1] // n_face
            n_slot = 12 if dim == 3 else 4
            n_ext = (24 if dim == 3 else 8) * m
            n_int = n_slot * m
            last = level == 1
            want_T = (not last) or return_T or subtree_recomp or return_h
            S = torch.empty((n_merges, n_int, n_ext), dtype=torch.float64, device=dev)
            g = torch.empty((n_merges, n_int, n_src), dtype=torch.float64, device=dev)
            T_out = torch.empty((n_merges, n_ext, n_ext), dtype=torch.float64, device=dev) if want_T else None
            h_out = torch.empty((n_merges, n_ext, n_src), dtype=torch.float64, device=dev) if want_T else None
            info = torch.zeros(n_merges, dtype=torch.int32, device=dev)
            # the C ABI takes at most MAX_BATCH merges per call (grid limits); wide low levels of deep 2D trees
            # are processed in contiguous slices, which also bounds the workspace
            step = min(n_merges, _lib.MAX_BATCH)
            need = ctypes.c_size_t()
            _lib.check(ws_fn(step, m, n_src, ctypes.byref(need)), "merge workspace query")
            ws = _lib.workspace(need.value, dev)
            logging.debug("merge level %d: %d merges, m=%d, workspace %.2f GB", level, n_merges, m, need.value / 2**30)
            def run_level():
                info.zero_()
                for s0 in range(0, n_merges, step):
                    s1 = min(n_merges, s0 + step)
                    rc = level_fn(_lib.stream_ptr(), s1 - s0, m, n_src, _lib.ptr(T[n_child * s0:n_child * s1]),
                                  _lib.ptr(h[n_child * s0:n_child * s1]), _lib.ptr(S[s0:s1]), _lib.ptr(g[s0:s1]),
                                  _lib.ptr(T_out[s0:s1]) if want_T else None, _lib.ptr(h_out[s0:s1]) if want_T else None,
                                  1 if want_T else 0, _lib.ptr(ws), ws.numel(), _lib.ptr(info[s0:s1]))
                    _lib.check(rc, "hps_merge_dtn_level")
                _lib.check_info(info, f"merge level {level}")

            _lib.with_pivoting_fallback(run_level)  # the merge reads T, h only: a repeated level starts from the same inputs
            del T, h
            T, h = T_out, h_out
            S_lst.append(S)
            g_lst.append(g if multi else g[..., 0])
        return S_lst, g_lst, T, (h if (h is None or multi) else h[..., 0])


def merge_stage_uniform_3D_DtN(T_arr, h_arr, l: int, device=None, host_device=None, return_T: bool = False):
    """Oct merges from the leaves to the root.  Returns ``(S_lst, g_tilde_lst[, T_last])``; the
    lists run from the level above the leaves to the root and, as in the reference, the root
    entries carry no batch axis (`merge/_uniform_3D_DtN.py:95-124`)."""
    S_lst, g_lst, T_last, _ = _merge_stage(T_arr, h_arr, l, 3, device, host_device, return_T, False)
    S_lst[-1] = S_lst[-1][0]
    g_lst[-1] = g_lst[-1][0]
    S_out = [_lib.to_result(S, host_device) for S in S_lst]
    g_out = [_lib.to_result(g, host_device) for g in g_lst]
    if return_T:
        return S_out, g_out, _lib.to_result(T_last[0], host_device)
    return S_out, g_out


def merge_stage_uniform_2D_DtN(T_arr, h_arr, l: int, device=None, host_device=None, subtree_recomp: bool = False,
                               return_T: bool = False):
    """Quad merges.  2D keeps a leading batch axis of 1 on the root entries
    (`merge/_uniform_2D_DtN.py:174-203`).  ``subtree_recomp=True`` returns only
    ``(T_last, h_last)`` with a leading axis of 1."""
    S_lst, g_lst, T_last, h_last = _merge_stage(T_arr, h_arr, l, 2, device, host_device, return_T, subtree_recomp)
    if subtree_recomp:
        return _lib.to_result(T_last, host_device), _lib.to_result(h_last, host_device)
    S_out = [_lib.to_result(S, host_device) for S in S_lst]
    g_out = [_lib.to_result(g, host_device) for g in g_lst]
    if return_T:
        return S_out, g_out, _lib.to_result(T_last[0], host_device)
    return S_out, g_out


def merge_subtrees_3D_DtN(T_arr, h_arr, l: int, n_roots: int, device=None):
    """Merge ``n_roots`` octree subtrees of depth ``l`` held back to back; everything stays on the
    device.  Returns ``(S_lst, g_tilde_lst, T_roots, h_roots)`` — the per-GPU part of the sharded
    build (the serial analogue is `_subtree_recomp.py:310-367`)."""
    S_lst, g_lst, T_roots, h_roots = _merge_stage(T_arr, h_arr, l, 3, device, device, True, False, n_roots=n_roots)
    return S_lst, g_lst, T_roots, h_roots


def merge_root_columns_3D_DtN(T8, h8, col0: int, ncols: int, device=None):
    """One rank's share of the root merge: ``S[:, col0:col0+ncols]`` and the full ``g_tilde`` from
    the 8 subtree-root operators (``hps_merge_oct_dtn_root_cols``)."""
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        T = _lib.to_device(T8, dev)
        h = _lib.to_device(h8, dev)
        multi = h.ndim == 3
        if not multi:
            h = h.unsqueeze(-1)
        n_src = h.shape[-1]
        m = T.shape[-1] // 6
        n_int = 12 * m
        S = torch.empty((n_int, ncols), dtype=torch.float64, device=dev)
        g = torch.empty((n_int, n_src), dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        need = ctypes.c_size_t()
        _lib.check(lib.hps_merge_oct_dtn_level_workspace(1, m, n_src, ctypes.byref(need)), "merge workspace query")
        ws = _lib.workspace(need.value, dev)
        def run():
            rc = lib.hps_merge_oct_dtn_root_cols(_lib.stream_ptr(), m, n_src, _lib.ptr(T), _lib.ptr(h), col0, ncols,
                                                 _lib.ptr(S), _lib.ptr(g), _lib.ptr(ws), ws.numel(), _lib.ptr(info))
            _lib.check(rc, "hps_merge_oct_dtn_root_cols")
            _lib.check_info(info, "root merge")

        _lib.with_pivoting_fallback(run)
        return S, (g if multi else g[..., 0])


def merge_stage_uniform_2D_ItI(T_arr, h_arr, l: int, device=None, host_device=None, subtree_recomp: bool = False,
                               return_T: bool = False, return_h: bool = False):
    """ItI quad merges, complex128 (reference `merge/_uniform_2D_ItI.py:19-179`).  Returns
    ``(S_lst, g_tilde_lst[, T_last][, h_last])``; every list entry keeps its batch axis (root: 1)."""
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        T = _lib.to_device(T_arr, dev, dtype=torch.complex128)
        h = _lib.to_device(h_arr, dev, dtype=torch.complex128)
        if T.ndim == 4:  # the reference's (n/4, 4, n_ext, n_ext) carry
            T = T.reshape(-1, T.shape[-2], T.shape[-1])
            h = h.reshape(T.shape[0], T.shape[-1], -1) if h.ndim == 4 else h.reshape(T.shape[0], T.shape[-1])
        multi = h.ndim == 3
        if not multi:
            h = h.unsqueeze(-1)
        n_src = h.shape[-1]
        if l <= 0 and T.shape[0] == 4:
            l = 1  # as in the reference, l = 0 on four operators still performs the final merge
        if T.shape[0] != 4**l:
            raise ValueError(f"expected {4**l} leaf operators for l={l}, got {T.shape[0]}")
        S_lst, g_lst = [], []
        c128 = dict(dtype=torch.complex128, device=dev)
        for level in range(l, 0, -1):
            n_merges = T.shape[0] // 4
            m = T.shape[-1] // 4
            n = 8 * m
            want_T = (level > 1) or return_T or subtree_recomp or return_h
            S = torch.empty((n_merges, n, n), **c128)
            g = torch.empty((n_merges, n, n_src), **c128)
            T_out = torch.empty((n_merges, n, n), **c128) if want_T else None
            h_out = torch.empty((n_merges, n, n_src), **c128) if want_T else None
            info = torch.zeros(n_merges, dtype=torch.int32, device=dev)
            step = min(n_merges, _lib.MAX_BATCH)
            need = ctypes.c_size_t()
            _lib.check(lib.hps_merge_quad_iti_level_workspace(step, m, n_src, ctypes.byref(need)), "workspace query")
            ws = _lib.workspace(need.value, dev)
            for s0 in range(0, n_merges, step):
                s1 = min(n_merges, s0 + step)
                rc = lib.hps_merge_quad_iti_level(_lib.stream_ptr(), s1 - s0, m, n_src, _lib.ptr(T[4 * s0:4 * s1]),
                                                  _lib.ptr(h[4 * s0:4 * s1]), _lib.ptr(S[s0:s1]), _lib.ptr(g[s0:s1]),
                                                  _lib.ptr(T_out[s0:s1]) if want_T else None,
                                                  _lib.ptr(h_out[s0:s1]) if want_T else None, 1 if want_T else 0,
                                                  _lib.ptr(ws), ws.numel(), _lib.ptr(info[s0:s1]))
                _lib.check(rc, "hps_merge_quad_iti_level")
            _lib.check_info(info, f"ItI merge level {level}")
            T, h = T_out, h_out
            S_lst.append(S)
            g_lst.append(g if multi else g[..., 0])
        if subtree_recomp:
            return _lib.to_result(T, host_device), _lib.to_result(h if multi else h[..., 0], host_device)
        out = ([_lib.to_result(S, host_device) for S in S_lst], [_lib.to_result(g, host_device) for g in g_lst])
        if return_T:
            out = out + (_lib.to_result(T[0], host_device),)
        if return_h:
            out = out + (_lib.to_result((h if multi else h[..., 0])[0], host_device),)
        return out



def __getattr__(name):  # same public names as the reference's `jaxhps.merge` package; imported lazily (those modules import this one)
    if name in ('merge_stage_adaptive_2D_DtN', 'merge_stage_adaptive_3D_DtN'):
        from . import adaptive

        return getattr(adaptive, name)
    if name in ('nosource_merge_stage_uniform_2D_DtN', 'nosource_merge_stage_uniform_2D_ItI'):
        from . import up_pass

        return getattr(up_pass, name)
    raise AttributeError(name)
