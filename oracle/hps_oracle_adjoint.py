"""TEST INFRASTRUCTURE (not product): derivatives of the 2D uniform HPS solve on the CPU, for checking
``jaxhps_b200/adjoint.py``.

The reference differentiates ``build_solver`` + ``solve`` with ``jax.jvp`` / ``jax.vjp``
(`/root/reference/examples/inverse_scattering_utils.py:110-171`, `check_autodiff_Jvp.py`, `check_autodiff_vJp.py`).
JAX is absent from the image and the NumPy shim cannot trace, so **parity with the reference's autodiff output is
unpinned**; what pins this file instead is self-consistency of the oracle: the closed-form tangent below is compared with
central finite differences of the (reference-pinned) oracle build + solve (`tests/test_oracle_adjoint.py`), and the
adjoint with the tangent through the dot-product identity.

* ``solve_full(pb, f, g)``      — no-source build (oracle) + up pass + down pass, everything recomputed from ``pb``'s
                                  coefficient fields (so that finite differences in the coefficients are possible);
* ``jvp_identity(...)``         — ``du = solve(df - sum_k dc_k (D_k u), dg)`` with the operators of the unperturbed build;
* ``jvp_finite_difference(...)``— central difference of ``solve_full``;
* ``vjp_dense(...)``            — the Jacobian assembled column by column from ``jvp_identity`` on a small problem, transposed.
"""
from __future__ import annotations

import copy

import numpy as np

from . import hps_oracle as orc

_NAMES = ("D_xx", "D_xy", "D_yy", "D_x", "D_y", "I")


def _build(pb):
    if pb.use_ItI:
        Y, T, Phi = orc.nosource_local_solve_stage_uniform_2D_ItI(pb)
        S, Di, BDi = orc.nosource_merge_stage_uniform_2D_ItI(T, pb.domain.L)
    else:
        Y, T, Phi = orc.nosource_local_solve_stage_uniform_2D_DtN(pb)
        S, Di, BDi = orc.nosource_merge_stage_uniform_2D_DtN(T, pb.domain.L)
    pb.Y, pb.Phi, pb.S_lst, pb.D_inv_lst, pb.BD_inv_lst = Y, Phi, S, Di, BDi
    return pb


def _solve_built(pb, f, g):
    """f (n_leaves, p^2, n_src), g (n_bdry, n_src) -> u (n_leaves, p^2, n_src)."""
    if pb.use_ItI:
        v, gl = orc.up_pass_uniform_2D_ItI(f, pb)
        return orc.down_pass_uniform_2D_ItI(g, pb.S_lst, gl, pb.Y, v)
    v, gl = orc.up_pass_uniform_2D_DtN(f, pb)
    return orc.down_pass_uniform_2D_DtN(g, pb.S_lst, gl, pb.Y, v)


def solve_full(pb, f, g):
    return _solve_built(_build(copy.copy(pb)), f, g)


def diff_op(pb, name):
    return np.eye(pb.domain.p ** 2) if name == "I" else getattr(pb, name)


def jvp_identity(pb_built, u, df, dg, dcoef):
    src = np.zeros_like(u) if df is None else np.array(df, dtype=u.dtype)
    for key, dc in (dcoef or {}).items():
        name = key[: -len("_coefficients")]
        src = src - np.asarray(dc)[..., None] * np.einsum("ij,njs->nis", diff_op(pb_built, name), u)
    dg = np.zeros((pb_built.domain.boundary_points.shape[0], u.shape[-1]), dtype=u.dtype) if dg is None else dg
    return _solve_built(pb_built, src, dg)


def jvp_finite_difference(pb, f, g, df, dg, dcoef, eps=1e-6):
    out = []
    for sgn in (+1.0, -1.0):
        q = copy.copy(pb)
        for key, dc in (dcoef or {}).items():
            setattr(q, key, getattr(pb, key) + sgn * eps * np.asarray(dc))
        ff = f if df is None else f + sgn * eps * df
        gg = g if dg is None else g + sgn * eps * dg
        out.append(solve_full(q, ff, gg))
    return (out[0] - out[1]) / (2 * eps)


def vjp_dense(pb_built, u, w, names):
    """Cotangents of (source, boundary data, coefficient fields in ``names``) for the cotangent ``w`` of ``u`` by
    assembling J^T w one unit tangent at a time (tiny problems only).  Single source (trailing axis 1)."""
    n_leaves, n_c, _ = u.shape
    n_b = pb_built.domain.boundary_points.shape[0]
    dt = u.dtype
    f_bar = np.zeros((n_leaves, n_c), dtype=dt)
    for i in range(n_leaves):
        for j in range(n_c):
            e = np.zeros((n_leaves, n_c, 1), dtype=dt)
            e[i, j, 0] = 1
            f_bar[i, j] = np.sum(w * jvp_identity(pb_built, u, e, None, None))
    g_bar = np.zeros(n_b, dtype=dt)
    for k in range(n_b):
        e = np.zeros((n_b, 1), dtype=dt)
        e[k, 0] = 1
        g_bar[k] = np.sum(w * jvp_identity(pb_built, u, None, e, None))
    c_bar = {}
    for name in names:
        cb = np.zeros((n_leaves, n_c), dtype=dt)
        for i in range(n_leaves):
            for j in range(n_c):
                e = np.zeros((n_leaves, n_c))
                e[i, j] = 1
                cb[i, j] = np.sum(w * jvp_identity(pb_built, u, None, None, {f"{name}_coefficients": e}))
        c_bar[name] = cb
    return f_bar, g_bar, c_bar
