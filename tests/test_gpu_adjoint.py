"""GPU: derivatives of the 2D uniform solve (`jaxhps_b200/adjoint.py`) against the oracle's closed-form
tangent (itself checked against finite differences on the CPU, `tests/test_oracle_adjoint.py`), the dot-product
identity between tangent and adjoint, and a dense adjoint assembled by the oracle on a tiny problem."""
import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200 import adjoint
from oracle import hps_oracle_adjoint as oadj
from test_oracle_adjoint import _problem

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


@pytest.mark.parametrize("iti,p,q,L", [(False, 6, 4, 2), (True, 6, 4, 2), (False, 8, 6, 3), (True, 8, 6, 3)])
def test_tangent_matches_the_oracle(iti, p, q, L):
    pb, f, g, df, dg, dco = _problem(iti, p, q, L, 21)
    built = oadj._build(hps.PDEProblem(pb.domain, **{k: getattr(pb, k) for k in dco}, **(dict(use_ItI=True, eta=pb.eta) if iti else {})))
    uo = oadj._solve_built(built, f, g)
    duo = oadj.jvp_identity(built, uo, df, dg, dco)
    hps.build_solver(pb)
    u, du = adjoint.solve_jvp(pb, g, f, d_source=df, d_boundary_data=dg, d_coefficients=dco)
    assert u.shape == uo.shape and _rel(u, uo) < 1e-10
    assert du.shape == duo.shape and _rel(du, duo) < 1e-9
    # single-source calling convention (no trailing axis)
    u1, du1 = adjoint.solve_jvp(pb, g[:, 0], f[..., 0], d_source=df[..., 0], d_boundary_data=dg[:, 0], d_coefficients=dco)
    assert u1.shape == uo.shape[:2] and _rel(du1, duo[..., 0]) < 1e-9


@pytest.mark.parametrize("iti,p,q,L,nsrc", [(False, 6, 4, 2, 1), (True, 6, 4, 2, 1), (False, 8, 6, 3, 2), (True, 8, 6, 3, 3)])
def test_adjoint_and_tangent_satisfy_the_dot_product_identity(iti, p, q, L, nsrc):
    pb, f, g, df, dg, dco = _problem(iti, p, q, L, 31)
    rng = np.random.default_rng(7)
    if nsrc > 1:
        rep = lambda a: np.concatenate([a * (1 + 0.3 * s) for s in range(nsrc)], axis=-1)  # noqa: E731
        f, g, df, dg = rep(f), rep(g), rep(df), rep(dg)
    hps.build_solver(pb)
    u, du = adjoint.solve_jvp(pb, g, f, d_source=df, d_boundary_data=dg, d_coefficients=dco)
    w = rng.normal(size=u.shape) + (1j * rng.normal(size=u.shape) if iti else 0)
    bars = adjoint.solve_vjp(pb, u, w)
    lhs = np.sum(w * du)
    rhs = np.sum(bars["source"] * df) + np.sum(bars["boundary_data"] * dg) + sum(np.sum(bars[k] * v) for k, v in dco.items())
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs)), (lhs, rhs)
    nb = pb.P.shape[0]
    assert np.all(np.asarray(bars["source"])[:, :nb] == 0)  # the source enters through the interior rows only


def test_adjoint_matches_the_dense_oracle_adjoint():
    pb, f, g, df, dg, dco = _problem(False, 4, 2, 1, 41)
    built = oadj._build(hps.PDEProblem(pb.domain, **{k: getattr(pb, k) for k in dco}))
    uo = oadj._solve_built(built, f, g)
    rng = np.random.default_rng(1)
    w = rng.normal(size=uo.shape)
    f_bar, g_bar, c_bar = oadj.vjp_dense(built, uo, w, ["I", "D_xx"])
    hps.build_solver(pb)
    u = hps.solve(pb, g, source=f)
    bars = adjoint.solve_vjp(pb, u, w)
    assert _rel(np.asarray(bars["source"])[..., 0], f_bar) < 1e-9
    assert _rel(np.asarray(bars["boundary_data"])[..., 0], g_bar) < 1e-9
    assert _rel(bars["I_coefficients"], c_bar["I"]) < 1e-9 and _rel(bars["D_xx_coefficients"], c_bar["D_xx"]) < 1e-9


@pytest.mark.parametrize("iti", [False, True])
def test_top_operator_tangent_matches_finite_differences_of_the_oracle(iti):
    """d(T_top)/dc (the operator the reference hands to its BIE coupling) against a central difference of the oracle's
    no-source merge stage."""
    import copy

    from oracle import hps_oracle as orc

    pb, f, g, df, dg, dco = _problem(iti, 6, 4, 2, 51)
    d = {"I_coefficients": dco["I_coefficients"], "D_yy_coefficients": dco["D_yy_coefficients"]}

    def top(q):
        if iti:
            _, T, _ = orc.nosource_local_solve_stage_uniform_2D_ItI(q)
            return orc.nosource_merge_stage_uniform_2D_ItI(T, q.domain.L, return_T=True)[3]
        _, T, _ = orc.nosource_local_solve_stage_uniform_2D_DtN(q)
        return orc.nosource_merge_stage_uniform_2D_DtN(T, q.domain.L, return_T=True)[3]

    eps, out = 1e-6, []
    for sgn in (1.0, -1.0):
        q = copy.copy(pb)
        for k, v in d.items():
            setattr(q, k, getattr(pb, k) + sgn * eps * v)
        out.append(top(q))
    fd = (out[0] - out[1]) / (2 * eps)
    hps.build_solver(pb)
    dT = adjoint.top_T_jvp(pb, d, chunk=17)
    assert dT.shape == fd.shape and _rel(dT, fd) < 1e-6


def test_inverse_scattering_forward_model_tangent_matches_finite_differences():
    """The whole chain of the reference's inverse-scattering example (`examples/inverse_scattering_utils.py:110-171`):
    q -> (I coefficient, source) -> build (root ItI operator) -> DtN -> BIE coupling -> incoming impedance -> solve.
    Tangent from `adjoint.scattering_forward_jvp` against a central difference of the same chain run on the GPU."""
    from jaxhps_b200 import scattering as sc

    rng = np.random.default_rng(61)
    k = 4.0
    dom = hps.Domain(8, 6, hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0), 2)
    pts = dom.interior_points
    shp = pts[..., 0].shape
    n_b = dom.boundary_points.shape[0]
    S = 0.2 * (rng.normal(size=(n_b, n_b)) + 1j * rng.normal(size=(n_b, n_b))) / np.sqrt(n_b)
    D = 0.2 * (rng.normal(size=(n_b, n_b)) + 1j * rng.normal(size=(n_b, n_b))) / np.sqrt(n_b)
    dirs = np.array([0.4])
    uin = sc.get_uin(k, pts.reshape(-1, 2), dirs).reshape(shp)
    q0 = 0.5 * np.exp(-((pts[..., 0] - 0.2) ** 2 + (pts[..., 1] + 0.1) ** 2) / 0.15**2)
    dq = rng.normal(size=shp)

    def forward(q):
        pb = hps.PDEProblem(dom, D_xx_coefficients=np.ones(shp), D_yy_coefficients=np.ones(shp),
                            I_coefficients=k**2 * (1 + q), use_ItI=True, eta=k)
        R = hps.build_solver(pb, return_top_T=True)
        T = sc.get_DtN_from_ItI(R, k)
        imp = sc.get_scattering_uscat_impedance(S, D, T, dirs, dom.boundary_points, k, k)
        src = -(k**2) * q * uin
        return pb, R, src, hps.solve(pb, imp[:, 0], source=src)

    pb, R, src, u = forward(q0)
    u_j, du = adjoint.scattering_forward_jvp(pb, R, src, {"I_coefficients": k**2 * dq}, S, D, dirs, k, d_source=-(k**2) * dq * uin)
    assert _rel(u_j, u) < 1e-10
    eps = 1e-6
    fd = (forward(q0 + eps * dq)[3] - forward(q0 - eps * dq)[3]) / (2 * eps)
    assert du.shape == fd.shape and _rel(du, fd) < 1e-6
    # adjoint of the same chain: <u_bar, du> == <source_bar, d_source> + <I_bar, dI>
    w = rng.normal(size=u.shape) + 1j * rng.normal(size=u.shape)
    bars = adjoint.scattering_forward_vjp(pb, R, src, u, w, S, D, dirs, k)
    lhs = np.sum(w * du)
    rhs = np.sum(bars["source"] * (-(k**2) * dq * uin)) + np.sum(bars["I_coefficients"] * (k**2 * dq))
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs)), (lhs, rhs)
    # and of the top-level operator alone
    Rb = rng.normal(size=np.asarray(R).shape) + 1j * rng.normal(size=np.asarray(R).shape)
    dR = adjoint.top_T_jvp(pb, {"I_coefficients": k**2 * dq}, chunk=19)
    cb = adjoint.top_T_vjp(pb, Rb, chunk=23)
    assert abs(np.sum(Rb * dR) - np.sum(cb["I_coefficients"] * (k**2 * dq))) < 1e-9 * max(1.0, abs(np.sum(Rb * dR)))
