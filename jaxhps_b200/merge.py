"""merge_stage on the B200: API mirror of `src/jaxhps/merge/_uniform_3D_DtN.py:12-124` and
`merge/_uniform_2D_DtN.py:13-203`; one ``hps_merge_*_dtn_level`` call per tree level."""
from __future__ import annotations

import ctypes
import logging

import torch

from . import _lib


def _merge_stage(T_arr, h_arr, l: int, dim: int, device, host_device, return_T: bool, subtree_recomp: bool,
                 return_h: bool = False):
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
        if T.shape[0] != n_child**l:
            raise ValueError(f"expected {n_child**l} leaf operators for l={l}, got {T.shape[0]}")
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
            need = ctypes.c_size_t()
            _lib.check(ws_fn(n_merges, m, n_src, ctypes.byref(need)), "merge workspace query")
            ws = _lib.WORKSPACE.get(need.value, dev)
            logging.debug("merge level %d: %d merges, m=%d, workspace %.2f GB", level, n_merges, m, need.value / 2**30)
            rc = level_fn(_lib.stream_ptr(), n_merges, m, n_src, _lib.ptr(T), _lib.ptr(h), _lib.ptr(S), _lib.ptr(g),
                          _lib.ptr(T_out), _lib.ptr(h_out), 1 if want_T else 0, _lib.ptr(ws), ws.numel(), _lib.ptr(info))
            _lib.check(rc, "hps_merge_dtn_level")
            _lib.check_info(info, f"merge level {level}")
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
