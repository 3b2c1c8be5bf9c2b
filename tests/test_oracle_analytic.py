"""Known-answer tests restating the reference's analytic accuracy suite
(/root/reference/tests/test_accuracy/cases.py:45-132, test_local_solve_stage_accuracy.py:43-53,
test_single_merge_accuracy.py:41-56) plus the 3D analogue the reference lacks."""
import numpy as np

import jaxhps_b200 as hps
from oracle import hps_oracle as orc

ATOL = 1e-12  # the reference's tolerance for polynomial data


def _poly2d(x):  # u = x^2 - y^2, harmonic
    return x[..., 0] ** 2 - x[..., 1] ** 2


def _dn_poly2d(b, root):
    """Outward normal derivative of x^2-y^2 on the S,E,N,W sides."""
    n = b.shape[0] // 4
    out = np.empty(b.shape[0])
    out[:n] = 2 * b[:n, 1]  # S: -du/dy = 2y
    out[n : 2 * n] = 2 * b[n : 2 * n, 0]  # E: du/dx
    out[2 * n : 3 * n] = -2 * b[2 * n : 3 * n, 1]  # N: du/dy
    out[3 * n :] = -2 * b[3 * n :, 0]  # W: -du/dx
    return out


def test_leaf_DtN_polynomial_2D():
    root = hps.DiscretizationNode2D(-np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2)
    dom = hps.Domain(16, 14, root, 0)
    one = np.ones_like(dom.interior_points[..., 0])
    pb = hps.PDEProblem(dom, source=np.zeros_like(one), D_xx_coefficients=one, D_yy_coefficients=one)
    Y, T, v, h = orc.local_solve_stage_uniform_2D_DtN(pb)
    g = _poly2d(dom.boundary_points)
    assert np.abs(Y[0] @ g - _poly2d(dom.interior_points[0])).max() < ATOL
    assert np.abs(T[0] @ g - _dn_poly2d(dom.boundary_points, root)).max() < 1e-10
    assert np.abs(v).max() < ATOL and np.abs(h).max() < ATOL


def test_leaf_particular_solution_2D():
    """Laplace u = f with u = x^2 + y^2 (f = 4): v + Y g reproduces u."""
    root = hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0)
    dom = hps.Domain(12, 10, root, 0)
    one = np.ones_like(dom.interior_points[..., 0])
    pb = hps.PDEProblem(dom, source=4 * one, D_xx_coefficients=one, D_yy_coefficients=one)
    Y, T, v, h = orc.local_solve_stage_uniform_2D_DtN(pb)
    u = lambda x: x[..., 0] ** 2 + x[..., 1] ** 2  # noqa: E731
    assert np.abs(Y[0] @ u(dom.boundary_points) + v[0] - u(dom.interior_points[0])).max() < ATOL


def test_single_merge_polynomial_2D():
    root = hps.DiscretizationNode2D(-np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2)
    dom = hps.Domain(6, 4, root, 1)
    one = np.ones_like(dom.interior_points[..., 0])
    pb = hps.PDEProblem(dom, source=np.zeros_like(one), D_xx_coefficients=one, D_yy_coefficients=one)
    Y, T, v, h = orc.local_solve_stage_uniform_2D_DtN(pb)
    S, g, T_top = orc.merge_stage_uniform_2D_DtN(T, h, 1, return_T=True)
    bd = _poly2d(dom.boundary_points)
    assert np.abs(T_top @ bd - _dn_poly2d(dom.boundary_points, root)).max() < 1e-11
    u = orc.down_pass_uniform_2D_DtN(bd, S, g, Y, v)
    assert np.abs(u - _poly2d(dom.interior_points)).max() < ATOL


def test_build_and_solve_polynomial_3D():
    """3D: u = x^2 - y^2 + z, two levels, variable (but consistent) source."""
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    dom = hps.Domain(6, 4, root, 2)
    x = dom.interior_points
    one = np.ones_like(x[..., 0])
    # operator: u_xx + u_yy + 2 u_zz + x u_z ; for u = x^2 - y^2 + z^2 + z:  2 - 2 + 4 + x (2z + 1)
    u_fn = lambda y: y[..., 0] ** 2 - y[..., 1] ** 2 + y[..., 2] ** 2 + y[..., 2]  # noqa: E731
    f = 4 + x[..., 0] * (2 * x[..., 2] + 1)
    pb = hps.PDEProblem(dom, source=f, D_xx_coefficients=one, D_yy_coefficients=one, D_zz_coefficients=2 * one,
                        D_z_coefficients=x[..., 0])
    Y, T, v, h = orc.local_solve_stage_uniform_3D_DtN(pb)
    S, g = orc.merge_stage_uniform_3D_DtN(T, h, 2)
    u = orc.down_pass_uniform_3D_DtN(u_fn(dom.boundary_points), S, g, Y, v)
    assert np.abs(u - u_fn(x)).max() < 1e-11


def test_interface_agreement_3D():
    """Children sharing an interface receive identical data after one down-propagation
    (3D port of /root/reference/tests/test_down_pass/test_uniform_2D_DtN.py:107-138)."""
    rng = np.random.default_rng(5)
    m = 4
    S = rng.normal(size=(12 * m, 24 * m))
    out = orc.propagate_down_oct_DtN(S, rng.normal(size=24 * m), rng.normal(size=12 * m))
    f = lambda c, face: out[c, face * m : (face + 1) * m]  # noqa: E731
    # a|b across x, a|d across y, a|e across z
    assert np.array_equal(f(0, 1), f(1, 0)) and np.array_equal(f(0, 3), f(3, 2)) and np.array_equal(f(0, 4), f(4, 5))
    assert np.array_equal(f(6, 0), f(7, 1)) and np.array_equal(f(2, 4), f(6, 5))


def test_iti_helmholtz_and_complex_coefficient_cases():
    """reference test_single_merge_accuracy.py:297-421: ItI build + solve on 4 leaves reproduces the analytic
    solution of a variable-coefficient Helmholtz problem and of one with a COMPLEX coefficient (atol 1e-8)."""
    from _iti_cases import COMPLEX, HELMHOLTZ, problem

    for case in (HELMHOLTZ, COMPLEX):
        dom, pb, g_in = problem(case)
        Y, R, v, h = orc.local_solve_stage_uniform_2D_ItI(pb)
        S_lst, g_lst = orc.merge_stage_uniform_2D_ItI(R, h, 1)
        u = orc.down_pass_uniform_2D_ItI(g_in, S_lst, g_lst, Y, v)
        assert np.abs(u - case["u"](dom.interior_points)).max() < 1e-8
