"""One-dimensional quadrature building blocks (host side, NumPy FP64).

Restates the behaviour of the reference's ``jaxhps.quadrature`` package
(`src/jaxhps/quadrature/_discretization.py:13-116`, `_differentiation.py:9-47`,
`_interpolation.py:18-326`) without JAX.  These run once per ``PDEProblem`` and
only produce the small constant matrices (1-D differentiation matrix, ``P``, ``Q``)
that the CUDA hot path consumes, so they stay on the host.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "chebyshev_points",
    "chebyshev_weights",
    "gauss_points",
    "affine_transform",
    "differentiation_matrix_1D",
    "barycentric_lagrange_interpolation_matrix_1D",
    "barycentric_lagrange_interpolation_matrix_2D",
    "barycentric_lagrange_interpolation_matrix_3D",
    "meshgrid_to_lst_of_pts",
]

_EPS = np.finfo(np.float64).eps


def chebyshev_points(n: int) -> np.ndarray:
    """``n`` Chebyshev–Lobatto points on [-1, 1], left end first
    (`_discretization.py:13-37`)."""
    theta = np.pi * (np.arange(n, dtype=np.float64) / (n - 1))
    return np.cos(theta[::-1])


def gauss_points(n: int) -> np.ndarray:
    """``n`` Gauss–Legendre nodes on [-1, 1] (`_discretization.py:105-116`)."""
    return np.polynomial.legendre.leggauss(n)[0]


def affine_transform(pts: np.ndarray, ab) -> np.ndarray:
    """Map points on [-1, 1] to [a, b] (`_discretization.py:90-101`)."""
    a, b = ab
    return 0.5 * (b - a) * np.asarray(pts) + 0.5 * (a + b)


def chebyshev_weights(n: int, bounds) -> np.ndarray:
    """Clenshaw–Curtis weights for ``chebyshev_points(n)`` scaled to ``bounds``
    (`_discretization.py:41-86`); DCT of the even-moment vector done with an FFT."""
    a, b = bounds
    moments = 2.0 / np.concatenate([[1.0], 1.0 - np.arange(2, n, 2) ** 2])
    half = n // 2 if n % 2 else n // 2 + 1
    mirrored = np.concatenate([moments, moments[1:half][::-1]])
    w = np.fft.ifft(mirrored).real
    out = np.concatenate([w, [w[0] / 2]])
    out[0] = w[0] / 2
    return out * (b - a) / 2


def differentiation_matrix_1D(points: np.ndarray) -> np.ndarray:
    """Chebyshev spectral differentiation matrix on ``points`` (Trefethen ch. 6;
    `_differentiation.py:9-47`): off-diagonal ``(c_i/c_j)/(x_i-x_j)``, diagonal by
    the negative-row-sum trick."""
    x = np.asarray(points, dtype=np.float64)
    p = x.shape[0]
    c = np.ones(p)
    c[0] = c[-1] = 2.0
    c[1::2] *= -1.0
    dx = x[:, None] - x[None, :]
    d = np.outer(c, 1.0 / c) / (dx + np.eye(p))
    return d - np.diag(d.sum(axis=1))


def _bary_weights_inv(nodes: np.ndarray) -> np.ndarray:
    """``w_j = prod_{k != j} (x_k - x_j)`` up to the sign convention the reference
    uses (product down the columns of ``x[:,None]-x[None,:]`` with unit diagonal)."""
    diff = nodes[:, None] - nodes[None, :]
    np.fill_diagonal(diff, 1.0)
    return np.prod(diff, axis=0)


def _bary_factor(from_pts: np.ndarray, to_pts: np.ndarray, eps_guard: bool):
    """Pieces of one tensor factor of the interpolation matrix: the (n_from, n_to) distance
    table, the inverse barycentric weights and the per-target normalisation."""
    from_pts = np.asarray(from_pts, dtype=np.float64)
    to_pts = np.asarray(to_pts, dtype=np.float64)
    w = _bary_weights_inv(from_pts)
    dist = to_pts[None, :] - from_pts[:, None]
    if eps_guard:
        dist = np.where(dist == 0, _EPS, dist)
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = np.sum(1.0 / (w[:, None] * dist), axis=0)
    return dist, w, norm


def barycentric_lagrange_interpolation_matrix_1D(from_pts, to_pts) -> np.ndarray:
    """(n_to, n_from) barycentric Lagrange matrix; a target coinciding with a source
    gets an exact unit row (`_interpolation.py:18-114`)."""
    from_pts = np.asarray(from_pts, dtype=np.float64)
    to_pts = np.asarray(to_pts, dtype=np.float64)
    dist, w, norm = _bary_factor(from_pts, to_pts, eps_guard=False)
    with np.errstate(divide="ignore", invalid="ignore"):
        mat = 1.0 / (dist.T * w[None, :] * norm[:, None])
    hit = to_pts[:, None] == from_pts[None, :]
    rows = hit.any(axis=1)
    mat[rows] = hit[rows].astype(np.float64)
    return mat


def _factor_matrix(from_pts, to_pts):
    """(n_to, n_from) factor with the reference's multi-D convention: exact zeros in the
    distance table are replaced by machine epsilon rather than special-cased
    (`_interpolation.py:189-194, 301-303`)."""
    dist, w, norm = _bary_factor(from_pts, to_pts, eps_guard=True)
    return dist.T, w, norm


def barycentric_lagrange_interpolation_matrix_2D(from_pts_x, from_pts_y, to_pts_x, to_pts_y) -> np.ndarray:
    """Tensor-product interpolation, rows/cols in ``meshgrid(..., indexing="ij")`` order
    (`_interpolation.py:118-213`).  Every entry is evaluated as one reciprocal of the
    full product, the same grouping the reference uses, so the two agree to rounding."""
    dx, wx, nx = _factor_matrix(from_pts_x, to_pts_x)
    dy, wy, ny = _factor_matrix(from_pts_y, to_pts_y)
    # index order (i, j, k, l) = (to_x, to_y, from_x, from_y)
    den = (
        dx[:, None, :, None]
        * dy[None, :, None, :]
        * wx[None, None, :, None]
        * wy[None, None, None, :]
        * nx[:, None, None, None]
        * ny[None, :, None, None]
    )
    return (1.0 / den).reshape(dx.shape[0] * dy.shape[0], dx.shape[1] * dy.shape[1])


def barycentric_lagrange_interpolation_matrix_3D(
    from_pts_x, from_pts_y, from_pts_z, to_pts_x, to_pts_y, to_pts_z
) -> np.ndarray:
    """3-D tensor-product interpolation (`_interpolation.py:217-326`)."""
    dx, wx, nx = _factor_matrix(from_pts_x, to_pts_x)
    dy, wy, ny = _factor_matrix(from_pts_y, to_pts_y)
    dz, wz, nz = _factor_matrix(from_pts_z, to_pts_z)
    den = (
        dx[:, None, None, :, None, None]
        * dy[None, :, None, None, :, None]
        * dz[None, None, :, None, None, :]
        * wx[None, None, None, :, None, None]
        * wy[None, None, None, None, :, None]
        * wz[None, None, None, None, None, :]
        * nx[:, None, None, None, None, None]
        * ny[None, :, None, None, None, None]
        * nz[None, None, :, None, None, None]
    )
    return (1.0 / den).reshape(
        dx.shape[0] * dy.shape[0] * dz.shape[0], dx.shape[1] * dy.shape[1] * dz.shape[1]
    )


def meshgrid_to_lst_of_pts(X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """Stack two ``(n, n)`` meshgrid arrays into an ``(n*n, 2)`` point list
    (`quadrature/_utils.py:8-25`)."""
    return np.stack([np.asarray(X).reshape(-1), np.asarray(Y).reshape(-1)], axis=-1)
