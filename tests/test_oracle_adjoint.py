"""CPU: the closed-form tangent of the 2D uniform solve (the identity `jaxhps_b200/adjoint.py` is built on) against
central finite differences of the reference-pinned oracle build + solve."""
import numpy as np
import pytest

import jaxhps_b200 as hps
from oracle import hps_oracle_adjoint as oadj


def _problem(iti, p, q, L, seed):
    rng = np.random.default_rng(seed)
    k = 3.0
    dom = hps.Domain(p, q, hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0), L)
    shp = dom.interior_points[..., 0].shape
    co = {"D_xx_coefficients": 1 + 0.1 * rng.normal(size=shp), "D_yy_coefficients": 1 + 0.1 * rng.normal(size=shp),
          "D_x_coefficients": 0.3 * rng.normal(size=shp), "I_coefficients": k**2 * (1 + 0.2 * rng.normal(size=shp))}
    extra = dict(use_ItI=True, eta=k) if iti else {}
    pb = hps.PDEProblem(dom, **co, **extra)
    cplx = (lambda s: 1j * rng.normal(size=s)) if iti else (lambda s: 0)
    nb = dom.boundary_points.shape[0]
    f = rng.normal(size=shp + (1,)) + cplx(shp + (1,))
    g = rng.normal(size=(nb, 1)) + cplx((nb, 1))
    df = rng.normal(size=shp + (1,)) + cplx(shp + (1,))
    dg = rng.normal(size=(nb, 1)) + cplx((nb, 1))
    dco = {key: rng.normal(size=shp) for key in co}
    return pb, f, g, df, dg, dco


@pytest.mark.parametrize("iti", [False, True])
def test_tangent_identity_matches_finite_differences(iti):
    pb, f, g, df, dg, dco = _problem(iti, 6, 4, 2, 11)
    built = oadj._build(pb)
    u = oadj._solve_built(built, f, g)
    du = oadj.jvp_identity(built, u, df, dg, dco)
    fd = oadj.jvp_finite_difference(pb, f, g, df, dg, dco, eps=1e-6)
    assert np.abs(du - fd).max() / np.abs(fd).max() < 1e-6


def test_dense_adjoint_satisfies_the_dot_product_identity():
    pb, f, g, df, dg, dco = _problem(False, 4, 2, 1, 3)
    built = oadj._build(pb)
    u = oadj._solve_built(built, f, g)
    rng = np.random.default_rng(0)
    w = rng.normal(size=u.shape)
    f_bar, g_bar, c_bar = oadj.vjp_dense(built, u, w, ["I", "D_x"])
    d = {k: v for k, v in dco.items() if k in ("I_coefficients", "D_x_coefficients")}
    du = oadj.jvp_identity(built, u, df, dg, d)
    lhs = np.sum(w * du)
    rhs = np.sum(f_bar * df[..., 0]) + np.sum(g_bar * dg[:, 0]) + sum(np.sum(c_bar[k[:-13]] * v) for k, v in d.items())
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))
