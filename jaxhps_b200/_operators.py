"""Per-problem constant operators: spectral differentiation in leaf ordering, ``P``
(Gauss boundary -> Chebyshev boundary), ``Q`` (Chebyshev cloud -> outward normal
derivative on the Gauss boundary) and the ItI analogues ``G``/``QH``.

Behavioural restatement of `src/jaxhps/_precompute_operators_2D.py:16-300` and
`src/jaxhps/_precompute_operators_3D.py:18-210`.  One-time host pre-compute; the hot
path receives these as constant device matrices (plus the scaled 1-D matrix ``D1`` from
which the CUDA leaf-assembly kernel rebuilds rows of the p^d x p^d operators on the fly).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ._grid import (
    face_cheby_indices_3D,
    rearrange_indices_ext_int_2D,
    rearrange_indices_ext_int_3D,
)
from .quadrature import (
    affine_transform,
    barycentric_lagrange_interpolation_matrix_1D,
    barycentric_lagrange_interpolation_matrix_2D,
    barycentric_lagrange_interpolation_matrix_3D,
    chebyshev_points,
    differentiation_matrix_1D,
    gauss_points,
)


def _permuted(M: np.ndarray, r: np.ndarray) -> np.ndarray:
    return M[np.ix_(r, r)]


def scaled_diff_matrix_1D(p: int, half_side_len: float) -> np.ndarray:
    """The 1-D matrix every leaf operator is a Kronecker product of."""
    return differentiation_matrix_1D(chebyshev_points(p)) / half_side_len


# ------------------------------------------------------------------ 2D


def precompute_diff_operators_2D(p: int, half_side_len: float) -> Tuple[np.ndarray, ...]:
    """Returns D_x, D_y, D_xx, D_yy, D_xy in boundary-first ordering.  y is stored
    descending, hence the sign on D_y (`_precompute_operators_2D.py:16-44`)."""
    r = rearrange_indices_ext_int_2D(p)
    D = scaled_diff_matrix_1D(p, half_side_len)
    eye = np.eye(p)
    dx = _permuted(np.kron(D, eye), r)
    dy = _permuted(-np.kron(eye, D), r)
    return dx, dy, dx @ dx, dy @ dy, dx @ dy


def precompute_P_2D_DtN(p: int, q: int) -> np.ndarray:
    """(4(p-1), 4q): each side's q Gauss values interpolated to that side's p Chebyshev
    points; the four corner rows average their two sides
    (`_precompute_operators_2D.py:49-79`)."""
    P1 = barycentric_lagrange_interpolation_matrix_1D(gauss_points(q), chebyshev_points(p))
    nb = 4 * (p - 1)
    out = np.zeros((nb, 4 * q))
    for side in range(4):
        rows = (side * (p - 1) + np.arange(p)) % nb
        out[rows, side * q : (side + 1) * q] += P1
    out[[0, p - 1, 2 * (p - 1), 3 * (p - 1)]] *= 0.5
    return out


def precompute_P_2D_ItI(p: int, q: int) -> np.ndarray:
    """(4(p-1), 4q) block diagonal of the 1-D map with its last row dropped
    (`_precompute_operators_2D.py:84-99`)."""
    P1 = barycentric_lagrange_interpolation_matrix_1D(gauss_points(q), chebyshev_points(p))
    return np.kron(np.eye(4), P1[:-1])


def precompute_N_matrix_2D(du_dx: np.ndarray, du_dy: np.ndarray, p: int) -> np.ndarray:
    """(4p, p^2) outward normal derivative at the 4p side points, corners counted on both
    of their sides (`_precompute_operators_2D.py:133-163`)."""
    nb = 4 * (p - 1)
    rows = [(side * (p - 1) + np.arange(p)) % nb for side in range(4)]
    return np.concatenate(
        [-du_dy[rows[0]], du_dx[rows[1]], du_dy[rows[2]], -du_dx[rows[3]]], axis=0
    )


def precompute_N_tilde_matrix_2D(du_dx: np.ndarray, du_dy: np.ndarray, p: int) -> np.ndarray:
    """(4(p-1), p^2) outward normal derivative at the 4(p-1) distinct boundary points, each
    corner assigned to the side that starts at it (`_precompute_operators_2D.py:167-192`)."""
    m = p - 1
    return np.concatenate(
        [-du_dy[:m], du_dx[m : 2 * m], du_dy[2 * m : 3 * m], -du_dx[3 * m : 4 * m]], axis=0
    )


def precompute_Q_2D_DtN(p: int, q: int, du_dx: np.ndarray, du_dy: np.ndarray) -> np.ndarray:
    """(4q, p^2) (`_precompute_operators_2D.py:103-129`)."""
    Q1 = barycentric_lagrange_interpolation_matrix_1D(chebyshev_points(p), gauss_points(q))
    return np.kron(np.eye(4), Q1) @ precompute_N_matrix_2D(du_dx, du_dy, p)


def precompute_G_2D_ItI(N_tilde: np.ndarray, eta: float) -> np.ndarray:
    """``N_tilde + i*eta*[I 0]``: Chebyshev cloud -> incoming impedance data on the
    4(p-1) boundary points (`_precompute_operators_2D.py:244-265`)."""
    G = N_tilde.astype(np.complex128)
    nb = G.shape[0]
    G[np.arange(nb), np.arange(nb)] += 1j * eta
    return G


def precompute_QH_2D_ItI(N: np.ndarray, p: int, q: int, eta: float) -> np.ndarray:
    """(4q, p^2) complex: outgoing impedance data ``u_n - i*eta*u`` on the 4p side points,
    interpolated to the Gauss boundary (`_precompute_operators_2D.py:196-240`)."""
    nb = 4 * (p - 1)
    H = N.astype(np.complex128)
    for side in range(4):
        cols = (side * (p - 1) + np.arange(p)) % nb
        H[side * p + np.arange(p), cols] -= 1j * eta
    Q1 = barycentric_lagrange_interpolation_matrix_1D(chebyshev_points(p), gauss_points(q))
    return np.kron(np.eye(4), Q1) @ H


def precompute_projection_ops_2D(q: int) -> Tuple[np.ndarray, np.ndarray]:
    """Refine (2q x q) / coarsen (q x 2q) maps between one Gauss panel and two half panels
    (`_precompute_operators_2D.py:269-300`)."""
    g = gauss_points(q)
    fine = np.concatenate([affine_transform(g, (-1.0, 0.0)), affine_transform(g, (0.0, 1.0))])
    return (
        barycentric_lagrange_interpolation_matrix_1D(g, fine),
        barycentric_lagrange_interpolation_matrix_1D(fine, g),
    )


# ------------------------------------------------------------------ 3D


def precompute_diff_operators_3D(p: int, half_side_len: float) -> Tuple[np.ndarray, ...]:
    """Returns D_x, D_y, D_z, D_xx, D_yy, D_zz, D_xy, D_xz, D_yz in boundary-first
    ordering (`_precompute_operators_3D.py:18-61`)."""
    r = rearrange_indices_ext_int_3D(p)
    D = scaled_diff_matrix_1D(p, half_side_len)
    eye = np.eye(p)
    dx = _permuted(np.kron(D, np.kron(eye, eye)), r)
    dy = _permuted(np.kron(eye, np.kron(D, eye)), r)
    dz = _permuted(np.kron(eye, np.kron(eye, D)), r)
    return dx, dy, dz, dx @ dx, dy @ dy, dz @ dz, dx @ dy, dx @ dz, dy @ dz


def precompute_P_3D_DtN(p: int, q: int) -> np.ndarray:
    """(p^3-(p-2)^3, 6q^2): face-wise 2-D interpolation Gauss -> Chebyshev; rows of
    Chebyshev points shared by 2 faces (edges) or 3 faces (corners) are averaged
    (`_precompute_operators_3D.py:65-147`)."""
    g, c = gauss_points(q), chebyshev_points(p)
    P2 = barycentric_lagrange_interpolation_matrix_2D(g, g, c, c)
    nb = p**3 - (p - 2) ** 3
    faces = face_cheby_indices_3D(p)
    out = np.zeros((nb, 6 * q * q))
    mult = np.zeros(nb)
    for f in range(6):
        out[faces[f], f * q * q : (f + 1) * q * q] = P2
        mult[faces[f]] += 1.0
    return out / mult[:, None]


def precompute_Q_3D_DtN(p: int, q: int, du_dx, du_dy, du_dz) -> np.ndarray:
    """(6q^2, p^3): outward normal derivative rows on each face's p x p Chebyshev grid,
    interpolated to that face's q x q Gauss grid (`_precompute_operators_3D.py:151-210`)."""
    g, c = gauss_points(q), chebyshev_points(p)
    Q2 = barycentric_lagrange_interpolation_matrix_2D(c, c, g, g)
    faces = face_cheby_indices_3D(p)
    normal = [(-1.0, du_dx), (1.0, du_dx), (-1.0, du_dy), (1.0, du_dy), (-1.0, du_dz), (1.0, du_dz)]
    return np.concatenate([Q2 @ (s * D[faces[f]]) for f, (s, D) in enumerate(normal)], axis=0)


# ------------------------------------------------------------------ adaptive discretisations


def precompute_projection_ops_3D(q: int) -> Tuple[np.ndarray, np.ndarray]:
    """Refine (4q^2 x q^2) and coarsen (q^2 x 4q^2) maps between one Gauss face panel and its four
    quarter panels A,B,C,D (quad order SW,SE,NE,NW in the face's two free coordinates).  The
    coarsening map is *not* an interpolant of all four panels: each coarse Gauss point is evaluated
    from the polynomial of the quarter panel that contains it (`_precompute_operators_3D.py:323-409`)."""
    if q % 2:
        raise ValueError("q must be even.")
    g = gauss_points(q)
    lo, hi = affine_transform(g, (-1.0, 0.0)), affine_transform(g, (0.0, 1.0))
    quarter = ((lo, lo), (hi, lo), (hi, hi), (lo, hi))  # (first coordinate, second coordinate)
    L_4f1 = np.concatenate([barycentric_lagrange_interpolation_matrix_2D(g, g, a, b) for a, b in quarter], axis=0)
    h = q // 2
    i, j = np.divmod(np.arange(q * q), q)  # coarse point = (first coordinate index, second)
    L_1f4 = np.zeros((q * q, 4 * q * q))
    for k, (a, b) in enumerate(quarter):
        rows = ((i >= h) if a is hi else (i < h)) & ((j >= h) if b is hi else (j < h))
        ga, gb = (g[h:] if a is hi else g[:h]), (g[h:] if b is hi else g[:h])
        L_1f4[rows, k * q * q : (k + 1) * q * q] = barycentric_lagrange_interpolation_matrix_2D(a, b, ga, gb)
    return L_4f1, L_1f4


def precompute_L_4f1(p: int) -> np.ndarray:
    """(4p^2, p^2) interpolation from a leaf's Chebyshev cloud to the clouds of its four children
    a..d, all in the boundary-first leaf ordering (`_precompute_operators_2D.py:303-379`)."""
    c = chebyshev_points(p)
    fine = np.concatenate([affine_transform(c, (-1.0, 0.0)), affine_transform(c, (0.0, 1.0))])
    M = barycentric_lagrange_interpolation_matrix_2D(c, c, fine, fine)
    r = rearrange_indices_ext_int_2D(p)
    ix, iy = np.divmod(np.arange(4 * p * p), 2 * p)
    # the leaf grid runs north -> south in its fast index (`_grid_creation_2D.py:53-68`)
    blocks = [(ix < p) & (iy >= p), (ix >= p) & (iy >= p), (ix >= p) & (iy < p), (ix < p) & (iy < p)]
    rows = np.concatenate([np.flatnonzero(b)[r] for b in blocks])
    return M[rows][:, r]


def precompute_L_8f1(p: int) -> np.ndarray:
    """(8p^3, p^3) interpolation from a leaf's Chebyshev cloud to those of its eight children a..h
    (`_precompute_operators_3D.py:412-550`)."""
    c = chebyshev_points(p)
    fine = np.concatenate([affine_transform(c, (-1.0, 0.0)), affine_transform(c, (0.0, 1.0))])
    M = barycentric_lagrange_interpolation_matrix_3D(c, c, c, fine, fine, fine)
    r = rearrange_indices_ext_int_3D(p)
    ii = np.arange(8 * p**3)
    ix, iy, iz = ii // (4 * p * p), (ii // (2 * p)) % (2 * p), ii % (2 * p)
    blocks = []
    for zhi in (True, False):  # a..d sit at z+, e..h at z-
        for xhi, yhi in ((False, False), (True, False), (True, True), (False, True)):
            blocks.append(((ix >= p) == xhi) & ((iy >= p) == yhi) & ((iz >= p) == zhi))
    rows = np.concatenate([np.flatnonzero(b)[r] for b in blocks])
    return M[rows][:, r]
