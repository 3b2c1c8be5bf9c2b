"""Source given at solve time (2D uniform trees): no-source build stages and the upward pass.

API mirror of the reference's
`local_solve/_nosource_uniform_2D_{DtN,ItI}.py`, `merge/_nosource_uniform_2D_{DtN,ItI}.py` and
`up_pass/_uniform_2D_{DtN,ItI}.py`.  ``Phi`` is obtained from the ordinary leaf kernels by solving
for the identity as a block of sources (``Phi = A_ii^-1`` resp. ``B^-1[:, n_b:]``); the merges keep
``D^-1`` and ``B D^-1`` (computed by the same pivoted LU with the identity as an extra right-hand
side); the up pass is a handful of bandwidth-bound batched mat-vecs."""
from __future__ import annotations

import ctypes
from typing import List

import torch

from . import _lib
from .local_solve import _ORDER_2D, _constants, _gather_coeffs, _gather_coeffs_complex

# ItI unknown orders (child, interface): the reference solves in [a5,a8,c6,c7,b5,b6,d7,d8] and returns
# rows of S / g~ in [a5,b5,b6,c6,c7,d7,d8,a8] (`merge/_uniform_2D_ItI.py:357-373`)
_SOLVE = [(0, 5), (0, 8), (2, 6), (2, 7), (1, 5), (1, 6), (3, 7), (3, 8)]
_OUT = [(0, 5), (1, 5), (1, 6), (2, 6), (2, 7), (3, 7), (3, 8), (0, 8)]
_SOLVE_POS_OF_OUT = [_SOLVE.index(k) for k in _OUT]  # position in the solve order of out-unknown u


def _block_perm(order_positions, m: int, dev) -> torch.Tensor:
    """Index vector that lists, block by block, rows `order_positions[k]*m + t`."""
    idx = [torch.arange(p * m, (p + 1) * m, device=dev) for p in order_positions]
    return torch.cat(idx)


# ----------------------------------------------------------------------------- leaf stages


def _nosource_leaves(pde_problem, iti: bool, device, host_device):
    return _lib.with_pivoting_fallback(lambda: _nosource_leaves_once(pde_problem, iti, device, host_device))


def _nosource_leaves_once(pde_problem, iti: bool, device, host_device):
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    dom = pde_problem.domain
    p, q = dom.p, dom.q
    n_c, n_i = p * p, (p - 2) ** 2
    n_b, n_g = n_c - n_i, 4 * q
    with torch.cuda.device(dev):
        if iti:
            coeffs, coeffs_im, which = _gather_coeffs_complex(pde_problem, _ORDER_2D, dev)
        else:
            (coeffs, which), coeffs_im = _gather_coeffs(pde_problem, _ORDER_2D, dev), None
        n_leaves = coeffs.shape[1]
        cdt = torch.complex128 if iti else torch.float64
        if iti:
            cache = pde_problem.__dict__.setdefault("_device_constants", {})
            key = ("iti", str(dev))
            if key not in cache:
                cache[key] = (_lib.to_device(pde_problem.D1, dev), _lib.to_device(pde_problem.P, dev),
                              _lib.to_device(pde_problem.G, dev, dtype=torch.complex128),
                              _lib.to_device(pde_problem.QH, dev, dtype=torch.complex128))
            D1, P, G, QH = cache[key]
        else:
            D1, P, Q = _constants(pde_problem, dev)
        Y = torch.empty((n_leaves, n_c, n_g), dtype=cdt, device=dev)
        T = torch.empty((n_leaves, n_g, n_g), dtype=cdt, device=dev)
        Phi = torch.empty((n_leaves, n_i, n_i) if not iti else (n_leaves, n_c, n_i), dtype=cdt, device=dev)
        # identity block of sources on the interior rows
        eye_src = torch.zeros((n_c, n_i), dtype=cdt, device=dev)
        eye_src[n_b:, :] = torch.eye(n_i, dtype=cdt, device=dev)
        one = ctypes.c_size_t()
        wsq = lib.hps_local_solve_2d_iti_workspace if iti else None
        if iti:
            _lib.check(wsq(1, p, q, n_i, ctypes.byref(one)), "workspace query")
        else:
            _lib.check(lib.hps_local_solve_dtn_workspace(2, 1, p, q, n_i, ctypes.byref(one)), "workspace query")
        per_leaf = one.value + (n_c + n_g) * n_i * (16 if iti else 8) * 2
        chunk = int(max(1, min(n_leaves, (4 << 30) // max(1, per_leaf), 65535)))
        need = ctypes.c_size_t()
        if iti:
            _lib.check(wsq(chunk, p, q, n_i, ctypes.byref(need)), "workspace query")
        else:
            _lib.check(lib.hps_local_solve_dtn_workspace(2, chunk, p, q, n_i, ctypes.byref(need)), "workspace query")
        ws = _lib.WORKSPACE.get(need.value, dev)
        for s in range(0, n_leaves, chunk):
            e = min(n_leaves, s + chunk)
            k = e - s
            src = eye_src.unsqueeze(0).expand(k, n_c, n_i).contiguous()
            v = torch.empty((k, n_c, n_i), dtype=cdt, device=dev)
            h = torch.empty((k, n_g, n_i), dtype=cdt, device=dev)
            info = torch.zeros(k, dtype=torch.int32, device=dev)
            c_chunk = coeffs[:, s:e].contiguous()
            if iti:
                ci_chunk = None if coeffs_im is None else coeffs_im[:, s:e].contiguous()
                rc = lib.hps_local_solve_2d_iti(_lib.stream_ptr(), k, p, q, n_i, which, _lib.ptr(c_chunk), _lib.ptr(D1),
                                                _lib.ptr(P), _lib.ptr(G), _lib.ptr(QH), _lib.ptr(src), _lib.ptr(Y[s:e]),
                                                _lib.ptr(T[s:e]), _lib.ptr(v), _lib.ptr(h), _lib.ptr(ws), ws.numel(),
                                                _lib.ptr(info), _lib.ptr(ci_chunk))
            else:
                rc = lib.hps_local_solve_dtn(_lib.stream_ptr(), 2, k, p, q, n_i, which, _lib.ptr(c_chunk), _lib.ptr(D1),
                                             _lib.ptr(P), _lib.ptr(Q), _lib.ptr(src), _lib.ptr(Y[s:e]), _lib.ptr(T[s:e]),
                                             _lib.ptr(v), _lib.ptr(h), _lib.ptr(ws), ws.numel(), _lib.ptr(info))
            _lib.check(rc, "hps_local_solve (no-source)")
            _lib.check_info(info, "no-source local solve")
            Phi[s:e] = v if iti else v[:, n_b:, :]
        return tuple(_lib.to_result(t, host_device) for t in (Y, T, Phi))


def nosource_local_solve_stage_uniform_2D_DtN(pde_problem, device=None, host_device=None):
    """``(Y, T, Phi)`` with ``Phi = A_ii^-1`` of shape ``(n, (p-2)^2, (p-2)^2)``
    (reference `local_solve/_nosource_uniform_2D_DtN.py:9-73`)."""
    return _nosource_leaves(pde_problem, False, device, host_device)


def nosource_local_solve_stage_uniform_2D_ItI(pde_problem, device=None, host_device=None):
    """``(Y, R, Phi)`` with ``Phi = B^-1[:, n_b:]`` of shape ``(n, p^2, (p-2)^2)``
    (reference `local_solve/_nosource_uniform_2D_ItI.py:13-77`)."""
    return _nosource_leaves(pde_problem, True, device, host_device)


# ----------------------------------------------------------------------------- merge stages


def _nosource_merge(T_arr, l: int, iti: bool, device, host_device, return_T: bool):
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    cdt = torch.complex128 if iti else torch.float64
    with torch.cuda.device(dev):
        T = _lib.to_device(T_arr, dev, dtype=cdt)
        if T.ndim == 4:
            T = T.reshape(-1, T.shape[-2], T.shape[-1])
        if l <= 0 and T.shape[0] == 4:
            l = 1  # as in the reference, l = 0 on four operators still performs the final merge
        if T.shape[0] != 4**l:
            raise ValueError(f"expected {4**l} leaf operators for l={l}, got {T.shape[0]}")
        S_lst, Di_lst, BDi_lst = [], [], []
        for level in range(l, 0, -1):
            n_merges = T.shape[0] // 4
            m = T.shape[-1] // 4
            n_int = 8 * m if iti else 4 * m
            n_ext = 8 * m
            S = torch.empty((n_merges, n_int, n_ext), dtype=cdt, device=dev)
            T_out = torch.empty((n_merges, n_ext, n_ext), dtype=cdt, device=dev)
            D_inv = torch.empty((n_merges, n_int, n_int), dtype=cdt, device=dev)
            BD_inv = torch.empty((n_merges, n_ext, n_int), dtype=cdt, device=dev)
            scratch = torch.empty(n_merges * 64 * m, dtype=torch.float64, device=dev)
            info = torch.zeros(n_merges, dtype=torch.int32, device=dev)
            need = ctypes.c_size_t()
            if iti:
                _lib.check(lib.hps_merge_quad_iti_level_workspace(n_merges, m, 1, ctypes.byref(need)), "workspace query")
                fn = lib.hps_merge_quad_iti_level_nosource
            else:
                _lib.check(lib.hps_merge_quad_dtn_level_workspace(n_merges, m, 1, ctypes.byref(need)), "workspace query")
                fn = lib.hps_merge_quad_dtn_level_nosource
            ws = _lib.WORKSPACE.get(need.value, dev)
            def run_level(fn=fn, n_merges=n_merges, m=m, T=T, S=S, T_out=T_out, D_inv=D_inv, BD_inv=BD_inv, scratch=scratch,
                          ws=ws, info=info, level=level):
                rc = fn(_lib.stream_ptr(), n_merges, m, _lib.ptr(T), _lib.ptr(S), _lib.ptr(T_out), _lib.ptr(D_inv),
                        _lib.ptr(BD_inv), _lib.ptr(scratch), _lib.ptr(ws), ws.numel(), _lib.ptr(info))
                _lib.check(rc, "hps_merge_quad_level_nosource")
                _lib.check_info(info, f"no-source merge level {level}")

            _lib.with_pivoting_fallback(run_level)
            # convert to the reference's stored layout: B D^-1 rows in pre-roll order, ItI D^-1 in solve order
            BD_inv = torch.roll(BD_inv, shifts=m, dims=1)
            if iti:
                # reference index (solve position) -> library index (out position)
                out_pos_of_solve = [_OUT.index(k) for k in _SOLVE]
                perm = _block_perm(out_pos_of_solve, m, dev)
                D_inv = D_inv.index_select(1, perm).index_select(2, perm)
                BD_inv = BD_inv.index_select(2, perm)
            S_lst.append(S)
            Di_lst.append(D_inv.contiguous())
            BDi_lst.append(BD_inv.contiguous())
            T = T_out
        out = tuple([_lib.to_result(x, host_device) for x in lst] for lst in (S_lst, Di_lst, BDi_lst))
        if return_T:
            out = out + (_lib.to_result(T[0], host_device),)
        return out


def nosource_merge_stage_uniform_2D_DtN(T_arr, l: int, device=None, host_device=None, return_T: bool = False):
    """``(S_lst, D_inv_lst, BD_inv_lst[, T_last])`` (reference `merge/_nosource_uniform_2D_DtN.py:13-130`)."""
    return _nosource_merge(T_arr, l, False, device, host_device, return_T)


def nosource_merge_stage_uniform_2D_ItI(T_arr, l: int, device=None, host_device=None, return_T: bool = False):
    """(reference `merge/_nosource_uniform_2D_ItI.py:19-150`)."""
    return _nosource_merge(T_arr, l, True, device, host_device, return_T)


# ----------------------------------------------------------------------------- upward pass


def _gemm(lib, dev, iti, M, N, K, alpha, A, lda, sA, B, sB, beta, C, ldc, sC, batch):
    """C = alpha A B + beta C, batched; B contiguous K x N."""
    if iti:
        ws = torch.empty(batch * 4 * K * N, dtype=torch.float64, device=dev)
        rc = lib.hps_zgemm_strided_batched(_lib.stream_ptr(), M, N, K, alpha, _lib.ptr(A), lda, sA, _lib.ptr(B), sB, beta,
                                           _lib.ptr(C), ldc, sC, batch, _lib.ptr(ws))
    else:
        rc = lib.hps_dgemm_strided_batched(_lib.stream_ptr(), M, N, K, alpha, _lib.ptr(A), lda, sA, _lib.ptr(B), N, sB,
                                           beta, _lib.ptr(C), ldc, sC, batch)
    _lib.check(rc, "gemm (up pass)")


def _up_pass(source, pde_problem, iti: bool, device, host_device, return_h_last: bool):
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    cdt = torch.complex128 if iti else torch.float64
    dom = pde_problem.domain
    p, q = dom.p, dom.q
    n_c, n_i = p * p, (p - 2) ** 2
    n_b, n_g = n_c - n_i, 4 * q
    with torch.cuda.device(dev):
        src = _lib.to_device(source, dev, dtype=cdt)
        multi = src.ndim == 3
        if not multi:
            src = src.unsqueeze(-1)
        n_leaves, _, n_src = src.shape
        Phi = _lib.to_device(pde_problem.Phi, dev, dtype=cdt)
        f_int = src[:, n_b:, :].contiguous()
        v = torch.zeros((n_leaves, n_c, n_src), dtype=cdt, device=dev)
        if iti:
            QH = _lib.to_device(pde_problem.QH, dev, dtype=cdt)
            _gemm(lib, dev, True, n_c, n_src, n_i, 1.0, Phi, n_i, n_c * n_i, f_int, n_i * n_src, 0.0, v, n_src,
                  n_c * n_src, n_leaves)
            Qm = QH
        else:
            Qm = _lib.to_device(pde_problem.Q, dev)
            v_int = v[:, n_b:, :]  # rows n_b.. of every leaf: contiguous block with stride n_c*n_src
            rc = lib.hps_dgemm_strided_batched(_lib.stream_ptr(), n_i, n_src, n_i, 1.0, _lib.ptr(Phi), n_i, n_i * n_i,
                                               _lib.ptr(f_int), n_src, n_i * n_src, 0.0, v_int.data_ptr(), n_src,
                                               n_c * n_src, n_leaves)
            _lib.check(rc, "Phi f")
        h = torch.empty((n_leaves, n_g, n_src), dtype=cdt, device=dev)
        _gemm(lib, dev, iti, n_g, n_src, n_c, 1.0, Qm, n_c, 0, v, n_c * n_src, 0.0, h, n_src, n_g * n_src, n_leaves)
        g_lst: List = []
        pos8 = (ctypes.c_int * 8)(*_SOLVE_POS_OF_OUT)
        for D_inv_h, BD_inv_h in zip(pde_problem.D_inv_lst, pde_problem.BD_inv_lst):
            D_inv = _lib.to_device(D_inv_h, dev, dtype=cdt)
            BD_inv = _lib.to_device(BD_inv_h, dev, dtype=cdt)
            n_nodes, n_int, _ = D_inv.shape
            m = h.shape[1] // 4
            n_ext = 8 * m
            h_int = torch.empty((n_nodes, n_int, n_src), dtype=cdt, device=dev)
            h_new = torch.empty((n_nodes, n_ext, n_src), dtype=cdt, device=dev)
            if iti:
                rc = lib.hps_up_gather_quad_iti(_lib.stream_ptr(), n_nodes, m, n_src, _lib.ptr(h), _lib.ptr(h_int),
                                                _lib.ptr(h_new), 1, pos8)
            else:
                rc = lib.hps_up_gather_quad(_lib.stream_ptr(), n_nodes, m, n_src, _lib.ptr(h), _lib.ptr(h_int),
                                            _lib.ptr(h_new), 1)
            _lib.check(rc, "hps_up_gather_quad")
            g = torch.empty((n_nodes, n_int, n_src), dtype=cdt, device=dev)
            # g~ = -D^-1 h_int ; h = h_ext - (B D^-1) h_int, both in the reference's stored layout
            _gemm(lib, dev, iti, n_int, n_src, n_int, -1.0, D_inv, n_int, n_int * n_int, h_int, n_int * n_src, 0.0, g,
                  n_src, n_int * n_src, n_nodes)
            _gemm(lib, dev, iti, n_ext, n_src, n_int, -1.0, BD_inv, n_int, n_ext * n_int, h_int, n_int * n_src, 1.0, h_new,
                  n_src, n_ext * n_src, n_nodes)
            h = torch.roll(h_new, shifts=-m, dims=1).contiguous()
            if iti:
                g = g.index_select(1, _block_perm(_SOLVE_POS_OF_OUT, m, dev))
            g_lst.append(g)
        # the DtN up pass of the reference keeps the source axis; the ItI one squeezes single sources
        squeeze = iti and not multi
        v_out = v[..., 0] if squeeze else v
        g_out = [g[..., 0] if squeeze else g for g in g_lst]
        out = (_lib.to_result(v_out, host_device), [_lib.to_result(g, host_device) for g in g_out])
        if return_h_last:
            h_last = h[0, :, 0] if squeeze else h[0]
            out = out + (_lib.to_result(h_last, host_device),)
        return out


def up_pass_uniform_2D_DtN(source, pde_problem, device=None, host_device=None, return_h_last: bool = False):
    """``(v, g_tilde_lst[, h_last])`` for a source given at solve time
    (reference `up_pass/_uniform_2D_DtN.py:8-107`; outputs keep the source axis, as there)."""
    return _up_pass(source, pde_problem, False, device, host_device, return_h_last)


def up_pass_uniform_2D_ItI(source, pde_problem, device=None, host_device=None, return_h_last: bool = False):
    """(reference `up_pass/_uniform_2D_ItI.py:8-136`)."""
    return _up_pass(source, pde_problem, True, device, host_device, return_h_last)
