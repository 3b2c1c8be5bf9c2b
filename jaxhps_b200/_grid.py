"""Leaf point clouds and orderings for uniform quad/oct-trees (host side).

Behavioural restatement of `src/jaxhps/_grid_creation_2D.py:21-124,204-233` and
`src/jaxhps/_grid_creation_3D.py:14-231,421-458`.  The orderings defined here are the
contract between the host layer and the CUDA kernels:

* leaves are stored in depth-first sibling order (2D: SW,SE,NE,NW; 3D: a..h);
* every leaf's Chebyshev cloud is listed boundary-first;
* boundary Gauss points follow the side/face order of the root.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from .quadrature import affine_transform, chebyshev_points, gauss_points

# ------------------------------------------------------------------ subdivision


def quad_children_bounds(bounds: np.ndarray) -> np.ndarray:
    """(n,4) [xmin,xmax,ymin,ymax] -> (n,4,4) children SW,SE,NE,NW
    (`_grid_creation_2D.py:35-46`)."""
    b = np.asarray(bounds, dtype=np.float64)
    x0, x1, y0, y1 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    xm, ym = (x0 + x1) / 2, (y0 + y1) / 2
    lo_hi_x = {0: (x0, xm), 1: (xm, x1)}
    lo_hi_y = {0: (y0, ym), 1: (ym, y1)}
    order = [(0, 0), (1, 0), (1, 1), (0, 1)]  # (x half, y half)
    kids = [np.stack([*lo_hi_x[ix], *lo_hi_y[iy]], axis=-1) for ix, iy in order]
    return np.stack(kids, axis=1)


def oct_children_bounds(bounds: np.ndarray) -> np.ndarray:
    """(n,6) -> (n,8,6) children a..h: a,b,c,d are the z+ layer in SW,SE,NE,NW order,
    e,f,g,h the z- layer (`_grid_creation_3D.py:29-74`)."""
    b = np.asarray(bounds, dtype=np.float64)
    x0, x1, y0, y1, z0, z1 = (b[:, i] for i in range(6))
    xm, ym, zm = (x0 + x1) / 2, (y0 + y1) / 2, (z0 + z1) / 2
    hx = {0: (x0, xm), 1: (xm, x1)}
    hy = {0: (y0, ym), 1: (ym, y1)}
    hz = {0: (z0, zm), 1: (zm, z1)}
    order = [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1), (0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]
    kids = [np.stack([*hx[ix], *hy[iy], *hz[iz]], axis=-1) for ix, iy, iz in order]
    return np.stack(kids, axis=1)


def uniform_leaf_bounds_2D(root, L: int) -> np.ndarray:
    bounds = np.array([[root.xmin, root.xmax, root.ymin, root.ymax]], dtype=np.float64)
    for _ in range(L):
        bounds = quad_children_bounds(bounds).reshape(-1, 4)
    return bounds


def uniform_leaf_bounds_3D(root, L: int) -> np.ndarray:
    bounds = np.array(
        [[root.xmin, root.xmax, root.ymin, root.ymax, root.zmin, root.zmax]], dtype=np.float64
    )
    for _ in range(L):
        bounds = oct_children_bounds(bounds).reshape(-1, 6)
    return bounds


# ------------------------------------------------------------------ leaf orderings


@lru_cache(maxsize=None)
def rearrange_indices_ext_int_2D(p: int) -> np.ndarray:
    """Permutation taking the natural p*p ordering (x index slowest, y stored
    *descending*) to boundary-first ordering: 4(p-1) boundary points walking
    S -> E -> N -> W from the SW corner, then the interior in natural order
    (`_grid_creation_2D.py:204-233`)."""
    i, j = np.meshgrid(np.arange(p), np.arange(p), indexing="ij")
    nat = (i * p + j).astype(np.int64)
    south = nat[:, p - 1]  # y = ymin, x ascending (includes both corners)
    east = nat[p - 1, p - 2 :: -1]  # x = xmax, y ascending, SE corner excluded
    north = nat[p - 2 : 0 : -1, 0]  # y = ymax, x descending, both corners excluded
    nw = nat[0:1, 0]
    west = nat[0, 1 : p - 1]  # x = xmin, y descending
    interior = nat[1 : p - 1, 1 : p - 1].reshape(-1)
    return np.concatenate([south, east, north, nw, west, interior])


@lru_cache(maxsize=None)
def rearrange_indices_ext_int_3D(p: int) -> np.ndarray:
    """Permutation taking the natural p^3 ordering (x slowest, z fastest) to
    [x- face, x+ face, rest of y-, rest of y+, rest of z-, rest of z+, interior]
    (`_grid_creation_3D.py:421-458`)."""
    i, j, k = np.meshgrid(np.arange(p), np.arange(p), np.arange(p), indexing="ij")
    i, j, k = i.reshape(-1), j.reshape(-1), k.reshape(-1)
    group = np.full(p**3, 6, dtype=np.int64)
    # later assignments must not override earlier (higher-priority) faces
    for code, mask in (
        (5, k == p - 1),
        (4, k == 0),
        (3, j == p - 1),
        (2, j == 0),
        (1, i == p - 1),
        (0, i == 0),
    ):
        group[mask] = code
    return np.argsort(group, kind="stable")


@lru_cache(maxsize=None)
def face_cheby_indices_3D(p: int) -> np.ndarray:
    """(6, p*p) positions, in the boundary-first leaf ordering, of the Chebyshev points
    lying on faces 0..5 (x-,x+,y-,y+,z-,z+), each face listed as a p x p grid with the
    first remaining coordinate slowest.  Same content as `get_face_{1..6}_idxes`
    (`_precompute_operators_3D.py:221-320`), obtained by inverting the leaf permutation."""
    r = rearrange_indices_ext_int_3D(p)
    where = np.empty(p**3, dtype=np.int64)
    where[r] = np.arange(p**3)
    nat = np.arange(p**3).reshape(p, p, p)
    faces = [nat[0], nat[p - 1], nat[:, 0, :], nat[:, p - 1, :], nat[:, :, 0], nat[:, :, p - 1]]
    return np.stack([where[f.reshape(-1)] for f in faces])


# ------------------------------------------------------------------ point clouds


def bounds_to_cheby_points_2D(bounds: np.ndarray, p: int) -> np.ndarray:
    """(n,4) -> (n,p*p,2) boundary-first Chebyshev clouds (`_grid_creation_2D.py:53-68`)."""
    c = chebyshev_points(p)
    b = np.asarray(bounds, dtype=np.float64)
    xs = 0.5 * (b[:, 1:2] - b[:, 0:1]) * c[None, :] + 0.5 * (b[:, 0:1] + b[:, 1:2])
    ys = 0.5 * (b[:, 3:4] - b[:, 2:3]) * c[None, :] + 0.5 * (b[:, 2:3] + b[:, 3:4])
    ys = ys[:, ::-1]
    X = np.broadcast_to(xs[:, :, None], (b.shape[0], p, p))
    Y = np.broadcast_to(ys[:, None, :], (b.shape[0], p, p))
    pts = np.stack([X, Y], axis=-1).reshape(b.shape[0], p * p, 2)
    return pts[:, rearrange_indices_ext_int_2D(p)]


def bounds_to_cheby_points_3D(bounds: np.ndarray, p: int) -> np.ndarray:
    """(n,6) -> (n,p^3,3) boundary-first Chebyshev clouds (`_grid_creation_3D.py:81-111`)."""
    c = chebyshev_points(p)
    b = np.asarray(bounds, dtype=np.float64)
    n = b.shape[0]
    ax = [
        0.5 * (b[:, 2 * d + 1 : 2 * d + 2] - b[:, 2 * d : 2 * d + 1]) * c[None, :]
        + 0.5 * (b[:, 2 * d : 2 * d + 1] + b[:, 2 * d + 1 : 2 * d + 2])
        for d in range(3)
    ]
    X = np.broadcast_to(ax[0][:, :, None, None], (n, p, p, p))
    Y = np.broadcast_to(ax[1][:, None, :, None], (n, p, p, p))
    Z = np.broadcast_to(ax[2][:, None, None, :], (n, p, p, p))
    pts = np.stack([X, Y, Z], axis=-1).reshape(n, p**3, 3)
    return pts[:, rearrange_indices_ext_int_3D(p)]


def compute_interior_Chebyshev_points_uniform_2D(root, L: int, p: int) -> np.ndarray:
    return bounds_to_cheby_points_2D(uniform_leaf_bounds_2D(root, L), p)


def compute_interior_Chebyshev_points_uniform_3D(root, L: int, p: int) -> np.ndarray:
    return bounds_to_cheby_points_3D(uniform_leaf_bounds_3D(root, L), p)


def compute_boundary_Gauss_points_uniform_2D(root, L: int, q: int) -> np.ndarray:
    """(4*2^L*q, 2) Gauss points walking S,E,N,W counter-clockwise from the SW corner
    (`_grid_creation_2D.py:77-124`)."""
    g = gauss_points(q)
    n = 2**L
    xb = np.linspace(root.xmin, root.xmax, n + 1)
    yb = np.linspace(root.ymin, root.ymax, n + 1)
    xg = np.concatenate([affine_transform(g, xb[i : i + 2]) for i in range(n)])
    yg = np.concatenate([affine_transform(g, yb[i : i + 2]) for i in range(n)])
    m = xg.shape[0]
    return np.concatenate(
        [
            np.column_stack([xg, np.full(m, root.ymin)]),
            np.column_stack([np.full(m, root.xmax), yg]),
            np.column_stack([xg[::-1], np.full(m, root.ymax)]),
            np.column_stack([np.full(m, root.xmin), yg[::-1]]),
        ]
    )


def _gauss_panels_2D(bounds4: np.ndarray, q: int) -> np.ndarray:
    """(n,4) panel bounds -> (n*q*q, 2) tensor Gauss grids, first coordinate slowest."""
    g = gauss_points(q)
    b = np.asarray(bounds4, dtype=np.float64)
    u = 0.5 * (b[:, 1:2] - b[:, 0:1]) * g[None, :] + 0.5 * (b[:, 0:1] + b[:, 1:2])
    v = 0.5 * (b[:, 3:4] - b[:, 2:3]) * g[None, :] + 0.5 * (b[:, 2:3] + b[:, 3:4])
    U = np.broadcast_to(u[:, :, None], (b.shape[0], q, q))
    V = np.broadcast_to(v[:, None, :], (b.shape[0], q, q))
    return np.stack([U, V], axis=-1).reshape(-1, 2)


def compute_boundary_Gauss_points_uniform_3D(root, L: int, q: int) -> np.ndarray:
    """(6*4^L*q^2, 3) Gauss points on faces x-,x+,y-,y+,z-,z+; inside a face the 4^L
    panels follow the quad recursion order SW,SE,NE,NW in that face's two free
    coordinates (`_grid_creation_3D.py:119-194`)."""
    faces2d = {
        "yz": np.array([[root.ymin, root.ymax, root.zmin, root.zmax]], dtype=np.float64),
        "xz": np.array([[root.xmin, root.xmax, root.zmin, root.zmax]], dtype=np.float64),
        "xy": np.array([[root.xmin, root.xmax, root.ymin, root.ymax]], dtype=np.float64),
    }
    for _ in range(L):
        faces2d = {k: quad_children_bounds(v).reshape(-1, 4) for k, v in faces2d.items()}
    yz = _gauss_panels_2D(faces2d["yz"], q)
    xz = _gauss_panels_2D(faces2d["xz"], q)
    xy = _gauss_panels_2D(faces2d["xy"], q)
    n = yz.shape[0]
    col = lambda v: np.full(n, v, dtype=np.float64)  # noqa: E731
    return np.concatenate(
        [
            np.column_stack([col(root.xmin), yz]),
            np.column_stack([col(root.xmax), yz]),
            np.column_stack([xz[:, 0], col(root.ymin), xz[:, 1]]),
            np.column_stack([xz[:, 0], col(root.ymax), xz[:, 1]]),
            np.column_stack([xy, col(root.zmin)]),
            np.column_stack([xy, col(root.zmax)]),
        ]
    )


# ------------------------------------------------------------------ adaptive trees


def leaf_bounds(root) -> np.ndarray:
    """(n_leaves, 2d) ``[xmin, xmax, ymin, ymax(, zmin, zmax)]`` of the leaves in depth-first order."""
    from ._tree import _bounds, get_all_leaves

    return np.array([[v for lim in _bounds(leaf) for v in lim] for leaf in get_all_leaves(root)], dtype=np.float64)


def compute_interior_Chebyshev_points_adaptive_2D(root, p: int) -> np.ndarray:
    """(n_leaves, p^2, 2) (`_grid_creation_2D.py:128-137`)."""
    return bounds_to_cheby_points_2D(leaf_bounds(root), p)


def compute_interior_Chebyshev_points_adaptive_3D(root, p: int) -> np.ndarray:
    """(n_leaves, p^3, 3) (`_grid_creation_3D.py:238-253`)."""
    return bounds_to_cheby_points_3D(leaf_bounds(root), p)


def compute_boundary_Gauss_points_adaptive_2D(root, q: int) -> np.ndarray:
    """Gauss points of the leaf sides on the domain boundary, counter-clockwise from the SW corner
    (`_grid_creation_2D.py:142-200`)."""
    from ._tree import get_ordered_lst_of_boundary_nodes

    g = gauss_points(q)
    S, E, N, W = get_ordered_lst_of_boundary_nodes(root)
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)  # noqa: E731
    xs = cat([affine_transform(g, (n.xmin, n.xmax)) for n in S])
    ye = cat([affine_transform(g, (n.ymin, n.ymax)) for n in E])
    xn = cat([affine_transform(g, (n.xmax, n.xmin)) for n in N])
    yw = cat([affine_transform(g, (n.ymax, n.ymin)) for n in W])
    return np.concatenate(
        [
            np.column_stack([xs, np.full(xs.shape[0], float(root.ymin))]),
            np.column_stack([np.full(ye.shape[0], float(root.xmax)), ye]),
            np.column_stack([xn, np.full(xn.shape[0], float(root.ymax))]),
            np.column_stack([np.full(yw.shape[0], float(root.xmin)), yw]),
        ]
    )


def compute_boundary_Gauss_points_adaptive_3D(root, q: int) -> np.ndarray:
    """Gauss points of the leaf faces on the domain boundary, face by face (x-,x+,y-,y+,z-,z+), each
    leaf face a q x q tensor grid with its first free coordinate slowest (`_grid_creation_3D.py:256-373`)."""
    from ._tree import get_ordered_lst_of_boundary_nodes

    faces = get_ordered_lst_of_boundary_nodes(root)
    fixed = [root.xmin, root.xmax, root.ymin, root.ymax, root.zmin, root.zmax]
    out = []
    for f, leaves in enumerate(faces):
        ax = f // 2
        if ax == 0:
            b = [[n.ymin, n.ymax, n.zmin, n.zmax] for n in leaves]
        elif ax == 1:
            b = [[n.xmin, n.xmax, n.zmin, n.zmax] for n in leaves]
        else:
            b = [[n.xmin, n.xmax, n.ymin, n.ymax] for n in leaves]
        uv = _gauss_panels_2D(np.array(b, dtype=np.float64), q)
        w = np.full(uv.shape[0], float(fixed[f]))
        cols = [uv[:, 0], uv[:, 1]]
        cols.insert(ax, w)
        out.append(np.column_stack(cols))
    return np.concatenate(out)
