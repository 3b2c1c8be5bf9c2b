"""The reference's analytic ItI cases (tests/test_accuracy/cases.py:134-265 of the reference) restated in
NumPy: domain [-pi/2, pi/2]^2, eta = 1, incoming impedance data u_n + i eta u."""
import numpy as np

import jaxhps_b200 as hps

ETA = 1.0
COMPLEX_COEFF = 1.0 + 0.1j  # k^2 + i gamma


def _helmholtz_u(x):
    return np.exp(1j * np.pi * x[..., 0]) + np.exp(1j * np.pi * x[..., 1])


HELMHOLTZ = dict(  # cases.py:134-209: Delta u + q u = f with a smooth potential
    u=_helmholtz_u,
    dudx=lambda x: 1j * np.pi * np.exp(1j * np.pi * x[..., 0]),
    dudy=lambda x: 1j * np.pi * np.exp(1j * np.pi * x[..., 1]),
    I=lambda x: 1.0 + np.exp(-x[..., 0] ** 2 - x[..., 1] ** 2),
    source=lambda x: -(np.pi**2) * _helmholtz_u(x) + (1.0 + np.exp(-x[..., 0] ** 2 - x[..., 1] ** 2)) * _helmholtz_u(x),
)

COMPLEX = dict(  # cases.py:212-265: Delta u + (k^2 + i gamma) u = f, u = x^3 + 3 y^2
    u=lambda x: x[..., 0] ** 3 + 3 * x[..., 1] ** 2 + 0j,
    dudx=lambda x: 3 * x[..., 0] ** 2 + 0j,
    dudy=lambda x: 6 * x[..., 1] + 0j,
    I=lambda x: COMPLEX_COEFF * np.ones_like(x[..., 0]),
    source=lambda x: 6 * x[..., 0] + 6 + COMPLEX_COEFF * (x[..., 0] ** 3 + 3 * x[..., 1] ** 2),
)


def problem(case, p=16, q=14, L=1):
    h = np.pi / 2
    dom = hps.Domain(p, q, hps.DiscretizationNode2D(-h, h, -h, h), L)
    X = dom.interior_points
    one = np.ones(X.shape[:2])
    pb = hps.PDEProblem(dom, source=case["source"](X), D_xx_coefficients=one, D_yy_coefficients=one,
                        I_coefficients=case["I"](X), use_ItI=True, eta=ETA)
    b = dom.boundary_points
    n = b.shape[0] // 4
    dn = np.concatenate([-case["dudy"](b[:n]), case["dudx"](b[n : 2 * n]), case["dudy"](b[2 * n : 3 * n]), -case["dudx"](b[3 * n :])])
    return dom, pb, dn + 1j * ETA * case["u"](b)
