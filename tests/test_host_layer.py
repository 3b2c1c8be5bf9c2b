"""Host pre-compute: structural properties the reference's own operator tests assert
(/root/reference/tests/test_precompute_operators_3D.py:79-165, test_precompute_operators_2D.py,
tests/test_quadrature/*), restated on NumPy."""
import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200 import _grid, _operators, quadrature as quad


def test_chebyshev_points_and_diff_matrix():
    x = quad.chebyshev_points(9)
    assert x[0] == -1 and x[-1] == 1 and np.all(np.diff(x) > 0)
    D = quad.differentiation_matrix_1D(x)
    for k in range(1, 8):
        assert np.abs(D @ x**k - k * x ** (k - 1)).max() < 1e-11
    assert np.abs(D.sum(axis=1)).max() < 1e-12


def test_clenshaw_curtis_weights_integrate_polynomials():
    for n in (6, 9):
        w = quad.chebyshev_weights(n, (0.0, 2.0))
        x = quad.affine_transform(quad.chebyshev_points(n), (0.0, 2.0))
        assert abs(w.sum() - 2.0) < 1e-13 and abs(w @ x**3 - 4.0) < 1e-12


def test_barycentric_interpolation_exact_on_polynomials():
    c, g = quad.chebyshev_points(8), quad.gauss_points(6)
    M = quad.barycentric_lagrange_interpolation_matrix_1D(c, g)
    assert np.abs(M @ c**5 - g**5).max() < 1e-13
    M2 = quad.barycentric_lagrange_interpolation_matrix_2D(c, c, g, g)
    X, Y = np.meshgrid(c, c, indexing="ij")
    XG, YG = np.meshgrid(g, g, indexing="ij")
    assert np.abs(M2 @ (X**2 * Y**3).ravel() - (XG**2 * YG**3).ravel()).max() < 1e-13
    same = quad.barycentric_lagrange_interpolation_matrix_1D(c, c)
    assert np.array_equal(same, np.eye(8))


@pytest.mark.parametrize("p", [4, 5, 8])
def test_leaf_orderings_are_permutations(p):
    r3 = _grid.rearrange_indices_ext_int_3D(p)
    assert sorted(r3) == list(range(p**3))
    r2 = _grid.rearrange_indices_ext_int_2D(p)
    assert sorted(r2) == list(range(p**2))
    faces = _grid.face_cheby_indices_3D(p)
    nb = p**3 - (p - 2) ** 3
    assert faces.shape == (6, p * p) and faces.max() < nb


def test_3D_operators_differentiate_polynomials_and_P_Q_are_consistent():
    p, q = 8, 6
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    dom = hps.Domain(p, q, root, 0)
    x = dom.interior_points[0]
    zero = np.zeros(p**3)
    pb = hps.PDEProblem(dom, source=zero[None], D_xx_coefficients=zero[None] + 1)
    f = x[:, 0] ** 2 * x[:, 1] + x[:, 2] ** 3
    assert np.abs(pb.D_x @ f - 2 * x[:, 0] * x[:, 1]).max() < 1e-11
    assert np.abs(pb.D_zz @ f - 6 * x[:, 2]).max() < 1e-9
    assert np.abs(pb.D_xy @ f - 2 * x[:, 0]).max() < 1e-9
    # P: Gauss boundary -> Chebyshev boundary is exact on low-degree polynomials
    b = dom.boundary_points
    nb = p**3 - (p - 2) ** 3
    fb = b[:, 0] * b[:, 1] + b[:, 2] ** 2
    assert np.abs(pb.P @ fb - (x[:nb, 0] * x[:nb, 1] + x[:nb, 2] ** 2)).max() < 1e-12
    # Q: outward normal derivative on the six faces
    g = pb.Q @ f
    n = q * q
    assert np.abs(g[:n] + 2 * b[:n, 0] * b[:n, 1]).max() < 1e-10  # x- face: -df/dx
    assert np.abs(g[5 * n :] - 3 * b[5 * n :, 2] ** 2).max() < 1e-10  # z+ face: +df/dz


def test_2D_operators_and_ItI_matrices():
    p, q, eta = 8, 6, 2.5
    root = hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0)
    dom = hps.Domain(p, q, root, 0)
    x = dom.interior_points[0]
    z = np.zeros((1, p * p))
    pb = hps.PDEProblem(dom, source=z, D_xx_coefficients=z + 1, use_ItI=True, eta=eta)
    f = x[:, 0] ** 3 + x[:, 0] * x[:, 1] ** 2
    assert np.abs(pb.D_y @ f - 2 * x[:, 0] * x[:, 1]).max() < 1e-11
    assert pb.G.shape == (4 * (p - 1), p * p) and pb.QH.shape == (4 * q, p * p) and pb.P.shape == (4 * (p - 1), 4 * q)
    # QH f = (f_n - i eta f) on the Gauss boundary
    b = dom.boundary_points
    fb = b[:, 0] ** 3 + b[:, 0] * b[:, 1] ** 2
    fn_south = -(2 * b[:q, 0] * b[:q, 1])
    assert np.abs(pb.QH[:q] @ f - (fn_south - 1j * eta * fb[:q])).max() < 1e-10


def test_domain_shapes_and_leaf_order():
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    dom = hps.Domain(4, 2, root, 2)
    assert dom.n_leaves == 64 and dom.interior_points.shape == (64, 64, 3) and dom.boundary_points.shape == (6 * 16 * 4, 3)
    # first child of the recursion is 'a' = (x-, y-, z+)
    c = dom.interior_points[0].mean(axis=0)
    assert c[0] < 0.25 and c[1] < 0.25 and c[2] > 0.75


def test_pdeproblem_validation_matches_reference_errors():
    r3 = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    d3 = hps.Domain(4, 2, r3, 1)
    s3 = np.zeros((8, 64))
    with pytest.raises(NotImplementedError):
        hps.PDEProblem(d3, source=s3, D_xx_coefficients=s3, use_ItI=True, eta=1.0)
    with pytest.raises(ValueError):
        hps.PDEProblem(d3, D_xx_coefficients=s3)  # 3D needs a source
    with pytest.raises(ValueError):
        hps.PDEProblem(d3, source=s3, D_xx_coefficients=np.zeros((8, 63)))
    r2 = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    d2 = hps.Domain(4, 2, r2, 1)
    s2 = np.zeros((4, 16))
    with pytest.raises(ValueError):
        hps.PDEProblem(d2, source=s2, D_zz_coefficients=s2)
    with pytest.raises(ValueError):
        hps.PDEProblem(d2, source=s2, D_xx_coefficients=s2, use_ItI=True)


def test_interpolation_matches_reference_fixture():
    """Domain.interp_{to,from}_interior_points against outputs of the reference (tests/golden/make_golden.py)."""
    import os

    G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interp_reference.npz")))
    d2 = hps.Domain(6, 4, hps.DiscretizationNode2D(-1.0, 1.0, 0.0, 2.0), 2)
    v, t = d2.interp_from_interior_points(G["f2"], np.linspace(-1, 1, 7), np.linspace(0, 2, 5))
    assert np.abs(v - G["from2"]).max() < 1e-13 and np.array_equal(t, G["pts2"])
    to = d2.interp_to_interior_points(G["g2"], np.linspace(-1, 1, 9), np.linspace(0, 2, 8))
    assert np.abs(to - G["to2"]).max() < 1e-12
    d3 = hps.Domain(4, 2, hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, -1.0, 0.0), 1)
    v, t = d3.interp_from_interior_points(G["f3"], np.linspace(0, 1, 5), np.linspace(0, 1, 4), np.linspace(-1, 0, 3))
    assert np.abs(v - G["from3"]).max() < 1e-13 and np.array_equal(t, G["pts3"])
    to = d3.interp_to_interior_points(G["g3"], np.linspace(0, 1, 5), np.linspace(0, 1, 6), np.linspace(-1, 0, 7))
    assert np.abs(to - G["to3"]).max() < 1e-12


def test_interpolation_is_exact_on_polynomials():
    d3 = hps.Domain(6, 4, hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0), 1)
    x = d3.interior_points
    f = x[..., 0] ** 3 - 2 * x[..., 1] * x[..., 2] ** 2
    xs = np.linspace(0.05, 0.95, 4)
    vals, pts = d3.interp_from_interior_points(f, xs, xs, xs)
    exact = pts[..., 0] ** 3 - 2 * pts[..., 1] * pts[..., 2] ** 2
    assert np.abs(vals - exact.reshape(vals.shape)).max() < 1e-12
