"""Derivatives of ``u = solve(pde_problem, g, source=f)`` with respect to the source, the boundary data and the
coefficient fields, for 2D uniform problems built WITHOUT a source (DtN and ItI).

The reference obtains these from ``jax.jvp`` / ``jax.vjp`` through its stage functions
(`examples/inverse_scattering_utils.py:110-171`, `examples/check_autodiff_Jvp.py`, `check_autodiff_vJp.py`,
`docs/Examples.rst:60-132`).  An opaque CUDA library needs them written by hand; no JAX runtime is
involved.  Both follow from one fact about the discretisation: the HPS solution is the solution of ONE global linear
system ``M(c) u = rhs(f, g)`` in which the coefficient fields enter only through the interior collocation rows,
``(M u)_i = sum_k c_k(x_i) (D_k u)_i = f_i``.  Hence, exactly (not up to discretisation error):

* ``jvp``:  ``du = M^-1 (df - sum_k dc_k . (D_k u),  dg)`` — one more solve with a modified source;
* ``vjp``:  with ``lambda = G_f^T u_bar`` (the transposed solve with respect to the source),
  ``f_bar = lambda``, ``g_bar = G_g^T u_bar``, ``c_k_bar = -lambda . (D_k u)``.

``G_f^T`` and ``G_g^T`` are the down pass and the up pass run backwards with transposed operators: the stored
``Y, S, D^-1, B D^-1, Phi, Q`` go through ``hps_gemv_t_strided_batched``; the boundary-data plumbing between tree levels
(which child face feeds which interface, flips, the exterior roll) is a signed partial permutation that is READ OFF
the forward kernels (``hps_up_gather_quad[_iti]``, ``hps_down_quad[_iti]_level`` applied to unit vectors at m = 3) and
transposed with ``index_add_`` — so the adjoint cannot drift from the kernels' conventions.  Complex problems use the
plain transpose (no conjugation), like ``jax.vjp`` of a holomorphic map.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from ._solve import solve
from .up_pass import _SOLVE_POS_OF_OUT, _block_perm

_COEFFS_2D = ("D_xx", "D_xy", "D_yy", "D_x", "D_y", "I")


def _check(pde_problem):
    dom = pde_problem.domain
    if not (dom.bool_2D and dom.bool_uniform):
        raise NotImplementedError("derivatives are implemented for 2D uniform problems (the reference's differentiable path)")
    if getattr(pde_problem, "Phi", None) is None or not pde_problem.D_inv_lst:
        raise ValueError("build the solver without a source (PDEProblem(source=None)) so that Phi, D_inv_lst and BD_inv_lst exist")


def _cdt(pde_problem):
    return torch.complex128 if pde_problem.use_ItI else torch.float64


def _as3(x, dev, dt):
    t = _lib.to_device(x, dev, dtype=dt)
    return t if t.ndim == 3 else t.unsqueeze(-1)


def _diff_op(pde_problem, name: str, dev, dt) -> Optional[torch.Tensor]:
    """(p^2, p^2) matrix of D_k on a leaf (identity for the I coefficient), in the solve dtype."""
    cache = pde_problem.__dict__.setdefault("_adjoint_ops", {})
    key = (name, str(dev), str(dt))
    if key not in cache:
        n_c = pde_problem.domain.p ** 2
        M = torch.eye(n_c, dtype=torch.float64) if name == "I" else torch.from_numpy(getattr(pde_problem, name).copy())
        cache[key] = M.to(dt).to(dev).contiguous()
    return cache[key]


def apply_diff_operator(pde_problem, name: str, u3: torch.Tensor) -> torch.Tensor:
    """``D_k u`` leaf by leaf: u3 (n_leaves, p^2, n_src) -> same shape (``hps_{d,z}gemm_strided_batched`` with the
    operator shared by all leaves)."""
    dev, dt = u3.device, u3.dtype
    lib = _lib.load()
    D = _diff_op(pde_problem, name, dev, dt)
    n_leaves, n_c, n_src = u3.shape
    out = torch.empty_like(u3)
    if dt.is_complex:
        ws = torch.empty(n_leaves * 4 * n_c * n_src, dtype=torch.float64, device=dev)
        rc = lib.hps_zgemm_strided_batched(_lib.stream_ptr(), n_c, n_src, n_c, 1.0, _lib.ptr(D), n_c, 0, _lib.ptr(u3),
                                           n_c * n_src, 0.0, _lib.ptr(out), n_src, n_c * n_src, n_leaves, _lib.ptr(ws))
    else:
        rc = lib.hps_dgemm_strided_batched(_lib.stream_ptr(), n_c, n_src, n_c, 1.0, _lib.ptr(D), n_c, 0, _lib.ptr(u3), n_src,
                                           n_c * n_src, 0.0, _lib.ptr(out), n_src, n_c * n_src, n_leaves)
    _lib.check(rc, "D_k u")
    return out


def _present(pde_problem):
    return [n for n in _COEFFS_2D if getattr(pde_problem, f"{n}_coefficients", None) is not None]


# ------------------------------------------------------------------------------------------------ jvp


def solve_jvp(pde_problem, boundary_data, source, d_source=None, d_boundary_data=None, d_coefficients: Optional[Dict] = None,
              u=None, compute_device=None, host_device=None):
    """Directional derivative of ``u = solve(pde_problem, boundary_data, source=source)``.

    ``d_coefficients``: ``{"I_coefficients": dI, "D_xx_coefficients": dDxx, ...}`` (arrays of the coefficients' shape;
    only fields the problem was built with).  Returns ``(u, du)`` in the shape ``solve`` returns.  The tangent costs
    one up + down pass with the operators of the existing build — nothing is re-factored."""
    _check(pde_problem)
    dev = _lib.require_cuda(compute_device)
    dt = _cdt(pde_problem)
    with torch.cuda.device(dev):
        src0 = _as3(source, dev, dt)                      # (n_leaves, p^2, n_src)
        n_src = src0.shape[-1]
        bd0 = _lib.to_device(boundary_data, dev, dtype=dt)
        single = _lib.to_device(source, dev, dtype=dt).ndim == 2
        bd2 = bd0.reshape(bd0.shape[0], -1)
        if bd2.shape[1] != n_src:
            bd2 = bd2.expand(-1, n_src)
        bd2 = bd2.contiguous()

        def run(src3, bdy2):  # always with explicit source axes (the reference's DtN up pass keeps them)
            out = solve(pde_problem, bdy2, source=src3, compute_device=dev, host_device=dev)
            out = _lib.to_device(out, dev, dtype=dt)
            return out if out.ndim == 3 else out.unsqueeze(-1)

        u3 = run(src0, bd2) if u is None else _as3(u, dev, dt)
        src = torch.zeros_like(u3) if d_source is None else _as3(d_source, dev, dt).clone()
        for key, dc in (d_coefficients or {}).items():
            name = key[: -len("_coefficients")] if key.endswith("_coefficients") else key
            if name not in _present(pde_problem):
                raise ValueError(f"{key}: the problem was built without this coefficient field")
            dc3 = _lib.to_device(dc, dev, dtype=dt).unsqueeze(-1)
            src = src - dc3 * apply_diff_operator(pde_problem, name, u3)
        if d_boundary_data is None:
            dbd = torch.zeros_like(bd2)
        else:
            dbd = _lib.to_device(d_boundary_data, dev, dtype=dt)
            dbd = dbd.reshape(dbd.shape[0], -1)
            if dbd.shape[1] != n_src:
                dbd = dbd.expand(-1, n_src)
            dbd = dbd.contiguous()
        du3 = run(src.contiguous(), dbd)
        if single:
            u3, du3 = u3[..., 0], du3[..., 0]
        return _lib.to_result(u3, host_device), _lib.to_result(du3, host_device)


def top_T_jvp(pde_problem, d_coefficients: Dict, chunk: int = 256, compute_device=None, host_device=None):
    """Directional derivative of the top-level Poincare-Steklov operator (``build_solver(..., return_top_T=True)``: the
    root DtN map, or the root ItI map ``R`` that the reference converts with ``get_DtN_from_ItI`` and feeds to its
    boundary-integral coupling, `examples/wave_scattering_utils.py:31-49`) with respect to the coefficient fields.

    Column j of ``T`` is the outgoing data of the homogeneous solution ``u_j`` with boundary data ``e_j``; perturbing the
    coefficients adds the source ``-sum_k dc_k (D_k u_j)`` with zero boundary data, whose outgoing data is the up pass'
    ``h_last``.  So ``dT[:, j] = h_last(-sum_k dc_k D_k u_j)``: one multi-source down pass and one multi-source up pass
    per chunk of boundary unknowns, no re-factorisation."""
    from .up_pass import up_pass_uniform_2D_DtN, up_pass_uniform_2D_ItI

    _check(pde_problem)
    dev = _lib.require_cuda(compute_device)
    dt = _cdt(pde_problem)
    up = up_pass_uniform_2D_ItI if pde_problem.use_ItI else up_pass_uniform_2D_DtN
    dom = pde_problem.domain
    n_b = dom.boundary_points.shape[0]
    n_leaves, n_c = dom.interior_points.shape[0], dom.p ** 2
    with torch.cuda.device(dev):
        dT = torch.empty((n_b, n_b), dtype=dt, device=dev)
        dcs = {}
        for key, dc in d_coefficients.items():
            name = key[: -len("_coefficients")] if key.endswith("_coefficients") else key
            if name not in _present(pde_problem):
                raise ValueError(f"{key}: the problem was built without this coefficient field")
            dcs[name] = _lib.to_device(dc, dev, dtype=dt).unsqueeze(-1)
        for j0 in range(0, n_b, chunk):
            j1 = min(n_b, j0 + chunk)
            E = torch.zeros((n_b, j1 - j0), dtype=dt, device=dev)
            E[torch.arange(j0, j1, device=dev), torch.arange(j1 - j0, device=dev)] = 1
            zero = torch.zeros((n_leaves, n_c, j1 - j0), dtype=dt, device=dev)
            U = _lib.to_device(solve(pde_problem, E, source=zero, compute_device=dev, host_device=dev), dev, dtype=dt)
            src = torch.zeros_like(U)
            for name, dc3 in dcs.items():
                src = src - dc3 * apply_diff_operator(pde_problem, name, U)
            h_last = up(src.contiguous(), pde_problem, device=dev, host_device=dev, return_h_last=True)[2]
            dT[:, j0:j1] = _lib.to_device(h_last, dev, dtype=dt).reshape(n_b, j1 - j0)
        return _lib.to_result(dT, host_device)


# ------------------------------------------------------------------------------------------------ plumbing maps


class _BlockMap:
    """A linear map between block vectors whose m x m blocks are 0, +-I or +-J (J = order reversal), stored per output
    block as a list of (input block, reversed, coefficient).  ``expand(m)`` gives flat (out index, in index, coefficient)
    triples for a block length m; ``apply_T`` is the transposed action on (n_nodes, n_out, n_src) data."""

    def __init__(self, G: torch.Tensor, m: int):
        n_out, n_in = G.shape[0] // m, G.shape[1] // m
        self.n_out_blocks, self.n_in_blocks = n_out, n_in
        self.terms = []
        eye = torch.eye(m, dtype=G.dtype)
        rev = torch.flip(eye, dims=[1])
        for ob in range(n_out):
            for ib in range(n_in):
                blk = G[ob * m:(ob + 1) * m, ib * m:(ib + 1) * m]
                if not bool(torch.any(blk != 0)):
                    continue
                c = blk[0, 0] if blk[0, 0] != 0 else blk[0, m - 1]
                if torch.equal(blk, c * eye):
                    self.terms.append((ob, ib, False, complex(c) if G.dtype.is_complex else float(c)))
                elif torch.equal(blk, c * rev):
                    self.terms.append((ob, ib, True, complex(c) if G.dtype.is_complex else float(c)))
                else:
                    raise RuntimeError("boundary plumbing is not a signed block permutation (kernel convention changed?)")
        self._cache = {}

    def expand(self, m: int, dev):
        key = (m, str(dev))
        if key not in self._cache:
            t = torch.arange(m)
            oi, ii, cf = [], [], []
            for ob, ib, rv, c in self.terms:
                oi.append(ob * m + t)
                ii.append(ib * m + (m - 1 - t if rv else t))
                cf.append(torch.full((m,), c, dtype=torch.complex128 if isinstance(c, complex) else torch.float64))
            self._cache[key] = (torch.cat(oi).to(dev), torch.cat(ii).to(dev), torch.cat(cf).to(dev))
        return self._cache[key]

    def apply_T(self, out_bar: torch.Tensor, m: int) -> torch.Tensor:
        oi, ii, cf = self.expand(m, out_bar.device)
        n_nodes, _, n_src = out_bar.shape
        in_bar = torch.zeros((n_nodes, self.n_in_blocks * m, n_src), dtype=out_bar.dtype, device=out_bar.device)
        in_bar.index_add_(1, ii, out_bar.index_select(1, oi) * cf.to(out_bar.dtype).view(1, -1, 1))
        return in_bar


_MAPS = {}


def _plumbing(iti: bool, dev):
    """(up-gather map, down-scatter map) of one quad merge, read off the forward kernels at m = 3.
    up:   [h_int (n_int) ; h_ext (8m, pre-roll order)]  <-  the four children's h (4 x 4m)
    down: the four children's boundary data (4 x 4m)    <-  [g_int (n_int) ; g_ext (8m)]"""
    key = (iti, str(dev))
    if key in _MAPS:
        return _MAPS[key]
    lib = _lib.load()
    m = 3
    dt = torch.complex128 if iti else torch.float64
    n_int = 8 * m if iti else 4 * m
    n_ext = 8 * m
    # ---- up gather
    n_in = 16 * m
    h_in = torch.eye(n_in, dtype=dt, device=dev).reshape(4, 4 * m, n_in).contiguous()
    h_int = torch.zeros((1, n_int, n_in), dtype=dt, device=dev)
    h_ext = torch.zeros((1, n_ext, n_in), dtype=dt, device=dev)
    if iti:
        pos8 = (ctypes.c_int * 8)(*_SOLVE_POS_OF_OUT)
        rc = lib.hps_up_gather_quad_iti(_lib.stream_ptr(), 1, m, n_in, _lib.ptr(h_in), _lib.ptr(h_int), _lib.ptr(h_ext), 1, pos8)
    else:
        rc = lib.hps_up_gather_quad(_lib.stream_ptr(), 1, m, n_in, _lib.ptr(h_in), _lib.ptr(h_int), _lib.ptr(h_ext), 1)
    _lib.check(rc, "hps_up_gather_quad (map extraction)")
    G_up = torch.cat([h_int[0], h_ext[0]], dim=0).cpu()
    # ---- down scatter: S = 0, so the level kernel returns the scatter of (g_int = g~, g_ext)
    n_b = n_int + n_ext
    basis = torch.eye(n_b, dtype=dt, device=dev)
    gt = basis[:n_int].reshape(1, n_int, n_b).contiguous()
    g_ext = basis[n_int:].reshape(1, n_ext, n_b).contiguous()
    S0 = torch.zeros((1, n_int, n_ext), dtype=dt, device=dev)
    out = torch.zeros((4, 4 * m, n_b), dtype=dt, device=dev)
    if iti:
        ws = torch.empty(48 * m * n_b, dtype=torch.float64, device=dev)
        rc = lib.hps_down_quad_iti_level(_lib.stream_ptr(), 1, m, n_b, _lib.ptr(S0), _lib.ptr(g_ext), _lib.ptr(gt), _lib.ptr(out),
                                         _lib.ptr(ws))
    else:
        ws = torch.empty((1, n_int, n_b), dtype=dt, device=dev)
        rc = lib.hps_down_quad_level(_lib.stream_ptr(), 1, m, n_b, _lib.ptr(S0), _lib.ptr(g_ext), _lib.ptr(gt), _lib.ptr(out),
                                     _lib.ptr(ws))
    _lib.check(rc, "hps_down_quad_level (map extraction)")
    G_dn = out.reshape(16 * m, n_b).cpu()
    _MAPS[key] = (_BlockMap(G_up, m), _BlockMap(G_dn, m))
    return _MAPS[key]


def _gemv_t(A: torch.Tensor, X: torch.Tensor, sA_shared: bool = False) -> torch.Tensor:
    """out[b] = A[b]^T X[b]: A (batch, M, K) or shared (M, K); X (batch, M, n_src) -> (batch, K, n_src)."""
    lib = _lib.load()
    batch, M, n_src = X.shape
    K = A.shape[-1]
    out = torch.empty((batch, K, n_src), dtype=X.dtype, device=X.device)
    A = A.contiguous()
    X = X.contiguous()
    rc = lib.hps_gemv_t_strided_batched(_lib.stream_ptr(), M, K, n_src, 1.0, _lib.ptr(A), K, 0 if sA_shared else M * K, _lib.ptr(X),
                                        n_src, M * n_src, 0.0, _lib.ptr(out), n_src, K * n_src, batch, 1 if X.dtype.is_complex else 0)
    _lib.check(rc, "hps_gemv_t_strided_batched")
    return out


# ------------------------------------------------------------------------------------------------ vjp


def _up_pass_T(pde_problem, gt_bars, h_top_bar, v_bar_leaf, dev) -> torch.Tensor:
    """The up pass run backwards (root -> leaves) with transposed operators: cotangents ``gt_bars[level]`` of the
    particular interface data g~ of every level (or None), ``h_top_bar`` of the root's outgoing data ``h_last``
    ((n_bdry, n_src) or None) and ``v_bar_leaf`` of the leaves' particular solution v ((n_leaves, p^2, n_src) or None)
    -> cotangent of the source, (n_leaves, p^2, n_src), zero on the boundary rows."""
    iti = bool(pde_problem.use_ItI)
    dt = _cdt(pde_problem)
    dom = pde_problem.domain
    p = dom.p
    n_c, n_i = p * p, (p - 2) ** 2
    n_b = n_c - n_i
    up_map, _ = _plumbing(iti, dev)
    n_levels = len(pde_problem.D_inv_lst)
    n_src = next(x.shape[-1] for x in ([h_top_bar, v_bar_leaf] + list(gt_bars or [])) if x is not None)
    h_bar = None if h_top_bar is None else h_top_bar.reshape(1, -1, n_src)
    for level in range(n_levels - 1, -1, -1):
        D_inv = _lib.to_device(pde_problem.D_inv_lst[level], dev, dtype=dt)
        BD_inv = _lib.to_device(pde_problem.BD_inv_lst[level], dev, dtype=dt)
        n_nodes, n_int, _ = D_inv.shape
        n_ext = BD_inv.shape[1]
        m = n_ext // 8
        h_int_bar = torch.zeros((n_nodes, n_int, n_src), dtype=dt, device=dev)
        if gt_bars is not None:
            gi = gt_bars[level]
            if iti:  # forward: g~ = (-D^-1 h_int)[perm]  ->  the cotangent goes back to the solve order
                perm = _block_perm(_SOLVE_POS_OF_OUT, m, dev)
                gs = torch.zeros_like(gi)
                gs.index_copy_(1, perm, gi)
                gi = gs
            h_int_bar = -_gemv_t(D_inv, gi)
        if h_bar is not None:
            h_new_bar = torch.roll(h_bar, shifts=m, dims=1).contiguous()   # forward: h = roll(h_new, -m)
            h_int_bar = h_int_bar - _gemv_t(BD_inv, h_new_bar)
        else:
            h_new_bar = torch.zeros((n_nodes, n_ext, n_src), dtype=dt, device=dev)
        h_bar = up_map.apply_T(torch.cat([h_int_bar, h_new_bar], dim=1), m).reshape(n_nodes * 4, 4 * m, n_src)
    # ---- leaves: v_bar = (direct) + Q^T h_bar, f_bar = Phi^T v_bar
    Qm = _lib.to_device(pde_problem.QH if iti else pde_problem.Q, dev, dtype=dt)
    v_bar = _gemv_t(Qm, h_bar, sA_shared=True)
    if v_bar_leaf is not None:
        v_bar = v_bar + v_bar_leaf
    Phi = _lib.to_device(pde_problem.Phi, dev, dtype=dt)
    f_bar = torch.zeros_like(v_bar)
    if iti:
        f_bar[:, n_b:, :] = _gemv_t(Phi, v_bar)               # Phi (n, n_c, n_i)
    else:
        f_bar[:, n_b:, :] = _gemv_t(Phi, v_bar[:, n_b:, :].contiguous())
    return f_bar




def solve_vjp(pde_problem, u, u_bar, compute_device=None, host_device=None) -> Dict:
    """Cotangents of ``u = solve(pde_problem, g, source=f)``: for any tangents ``(df, dg, dc)``,
    ``sum(u_bar * du) == sum(source_bar * df) + sum(boundary_bar * dg) + sum_k sum(c_k_bar * dc_k)`` (bilinear pairing,
    no conjugation).  Returns ``{"source": ..., "boundary_data": ..., "<name>_coefficients": ...}``; with several sources
    the coefficient cotangents are summed over the source axis."""
    _check(pde_problem)
    dev = _lib.require_cuda(compute_device)
    iti = bool(pde_problem.use_ItI)
    dt = _cdt(pde_problem)
    dom = pde_problem.domain
    p, q = dom.p, dom.q
    n_c, n_i = p * p, (p - 2) ** 2
    n_b, n_g = n_c - n_i, 4 * q
    with torch.cuda.device(dev):
        u_t = _lib.to_device(u, dev, dtype=dt)
        single = u_t.ndim == 2
        u3 = u_t if not single else u_t.unsqueeze(-1)
        w = _as3(u_bar, dev, dt).contiguous()
        n_leaves, _, n_src = w.shape
        up_map, dn_map = _plumbing(iti, dev)
        Y = _lib.to_device(pde_problem.Y, dev, dtype=dt)
        S_lst = [_lib.to_device(S, dev, dtype=dt) for S in pde_problem.S_lst]
        # ---- transposed down pass: leaves -> root
        g_bar = _gemv_t(Y, w)                                   # (n_leaves, n_g, n_src)
        gt_bars = []
        for S in S_lst:
            n_nodes, n_int, n_ext = S.shape
            m = n_ext // 8
            kids = g_bar.reshape(n_nodes, 16 * m, n_src)          # four children x 4m
            both = dn_map.apply_T(kids, m)                        # [g_int_bar ; g_ext_bar]
            gi, ge = both[:, :n_int].contiguous(), both[:, n_int:]
            gt_bars.append(gi)
            g_bar = ge + _gemv_t(S, gi)
        boundary_bar = g_bar[0]                                   # (n_bdry, n_src)
        f_bar = _up_pass_T(pde_problem, gt_bars, None, w, dev)
        out = {"source": f_bar[..., 0] if single else f_bar, "boundary_data": boundary_bar[..., 0] if single else boundary_bar}
        for name in _present(pde_problem):
            out[f"{name}_coefficients"] = -(f_bar * apply_diff_operator(pde_problem, name, u3)).sum(dim=-1)
        return {k: _lib.to_result(v, host_device) for k, v in out.items()}


def scattering_forward_jvp(pde_problem, R_top, source, d_coefficients: Dict, S, D, source_dirs, k: float, d_source=None,
                           compute_device=None, host_device=None):
    """Tangent of the reference's inverse-scattering forward model (`examples/inverse_scattering_utils.py:110-171`):
    coefficients -> root ItI operator ``R`` -> ``T = get_DtN_from_ItI(R)`` -> incoming impedance data of the scattered
    field (BIE coupling with the layer potentials ``S``, ``D``) -> ``u = solve(pde_problem, imp, source=source)``.
    Returns ``(u, du)``.  ``R_top``: what ``build_solver(pde_problem, return_top_T=True)`` returned (no-source build)."""
    from . import scattering as sc

    _check(pde_problem)
    if not pde_problem.use_ItI:
        raise ValueError("the scattering coupling is formulated for ItI problems")
    dev = _lib.require_cuda(compute_device)
    eta = pde_problem.eta
    dR = top_T_jvp(pde_problem, d_coefficients, compute_device=dev, host_device=dev)
    T, dT = sc.get_DtN_from_ItI_jvp(R_top, dR, eta, device=dev, host_device=dev)
    imp, dimp = sc.get_scattering_uscat_impedance_jvp(S, D, T, dT, source_dirs, pde_problem.domain.boundary_points, k, eta,
                                                      device=dev, host_device=dev)
    return solve_jvp(pde_problem, imp, source, d_source=d_source, d_boundary_data=dimp, d_coefficients=d_coefficients,
                     compute_device=dev, host_device=host_device)


def top_T_vjp(pde_problem, T_bar, chunk: int = 256, compute_device=None, host_device=None) -> Dict:
    """Cotangents of the coefficient fields for a cotangent ``T_bar`` (n_bdry, n_bdry) of the top-level operator:
    ``sum(T_bar * dT) == sum_k sum(c_k_bar * dc_k)`` with ``dT = top_T_jvp(dc)``.  Column j of ``dT`` is the up pass'
    ``h_last`` for the source ``-sum_k dc_k D_k u_j``, so ``c_k_bar = -sum_j (H^T T_bar[:, j]) . (D_k u_j)`` with ``H^T`` the
    transposed up pass (:func:`_up_pass_T`) — per chunk of boundary unknowns one multi-source down pass and one transposed
    up pass."""
    _check(pde_problem)
    dev = _lib.require_cuda(compute_device)
    dt = _cdt(pde_problem)
    dom = pde_problem.domain
    n_b = dom.boundary_points.shape[0]
    n_leaves, n_c = dom.interior_points.shape[0], dom.p ** 2
    with torch.cuda.device(dev):
        Tb = _lib.to_device(T_bar, dev, dtype=dt)
        names = _present(pde_problem)
        bars = {name: torch.zeros((n_leaves, n_c), dtype=dt, device=dev) for name in names}
        for j0 in range(0, n_b, chunk):
            j1 = min(n_b, j0 + chunk)
            E = torch.zeros((n_b, j1 - j0), dtype=dt, device=dev)
            E[torch.arange(j0, j1, device=dev), torch.arange(j1 - j0, device=dev)] = 1
            zero = torch.zeros((n_leaves, n_c, j1 - j0), dtype=dt, device=dev)
            U = _lib.to_device(solve(pde_problem, E, source=zero, compute_device=dev, host_device=dev), dev, dtype=dt)
            lam = _up_pass_T(pde_problem, None, Tb[:, j0:j1].contiguous(), None, dev)
            for name in names:
                bars[name] -= (lam * apply_diff_operator(pde_problem, name, U)).sum(dim=-1)
        return {f"{k}_coefficients": _lib.to_result(v, host_device) for k, v in bars.items()}


def scattering_forward_vjp(pde_problem, R_top, source, u, u_bar, S, D, source_dirs, k: float, compute_device=None,
                           host_device=None) -> Dict:
    """Cotangents of ``(source, coefficient fields)`` for the reference's inverse-scattering forward model (see
    :func:`scattering_forward_jvp`): ``sum(u_bar * du) == sum(source_bar * d_source) + sum_k sum(c_k_bar * dc_k)``.
    The solve's own adjoint gives the direct part and the cotangent of the incoming impedance data; that is pulled back
    through the BIE coupling (two transposed dense solves), the ItI -> DtN conversion and the top-level operator."""
    from . import scattering as sc

    _check(pde_problem)
    dev = _lib.require_cuda(compute_device)
    eta = pde_problem.eta
    with torch.cuda.device(dev):
        bars = solve_vjp(pde_problem, u, u_bar, compute_device=dev, host_device=dev)
        imp_bar = bars.pop("boundary_data")
        imp_bar = imp_bar.reshape(imp_bar.shape[0], -1)
        T_bar = sc.get_scattering_uscat_impedance_vjp(S, D, sc.get_DtN_from_ItI(R_top, eta, device=dev, host_device=dev), imp_bar,
                                                      source_dirs, pde_problem.domain.boundary_points, k, eta, device=dev,
                                                      host_device=dev)
        R_bar = sc.get_DtN_from_ItI_vjp(R_top, T_bar, eta, device=dev, host_device=dev)
        top = top_T_vjp(pde_problem, R_bar, compute_device=dev, host_device=dev)
        out = {"source": bars["source"]}
        for key, v in top.items():
            out[key] = bars[key] + v
        return {k2: _lib.to_result(v, host_device) for k2, v in out.items()}
