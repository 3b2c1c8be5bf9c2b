"""Adaptive mesh generation: refine a quadtree / octree until a function is resolved, keeping
neighbouring leaves within one level of each other.

Restates `src/jaxhps/_adaptive_discretization_2D.py:37-328` and `_adaptive_discretization_3D.py:35-395`
(one implementation for both dimensions).  Host work: every round evaluates the refinement criterion
for the whole queue in one batched NumPy call (the reference vmaps the same check over the queue).

Criterion per queued box (`_adaptive_discretization_3D.py:466-503`): evaluate ``f`` on the box's own
Chebyshev cloud and on the clouds of its 2^d children, interpolate the coarse samples to the fine
clouds, and accept the box when

* L_inf:  max|interp - fine| / max(global_norm, max|fine|) < tol, the running global norm being
  raised after each round;
* L_2:    sum(w (interp - fine)^2) / ||f||_2^2 < tol^2 with Clenshaw-Curtis weights w and the
  squared norm estimated once on the first subdivision of the root.

Deviation: neighbour look-ups use the boxes' integer positions at their depth instead of comparing
floating-point corner coordinates, so trees over domains with inexact midpoints cannot miss a
neighbour by one ulp."""
from __future__ import annotations

import logging
from typing import Callable, List

import numpy as np

from ._grid import (
    bounds_to_cheby_points_2D,
    bounds_to_cheby_points_3D,
    oct_children_bounds,
    quad_children_bounds,
    rearrange_indices_ext_int_2D,
    rearrange_indices_ext_int_3D,
)
from ._operators import precompute_L_4f1, precompute_L_8f1
from ._tree import _OFFSETS_2D, _OFFSETS_3D, _add_children, _bounds, _is_2D, get_all_leaves
from .quadrature import chebyshev_weights

__all__ = [
    "generate_adaptive_mesh_level_restriction_2D",
    "generate_adaptive_mesh_level_restriction_3D",
    "get_squared_l2_norm_single_panel",
    "get_squared_l2_norm_single_voxel",
    "node_to_bounds",
]


def node_to_bounds(node) -> np.ndarray:
    """``[xmin, xmax, ymin, ymax(, zmin, zmax)]`` of one node (`_adaptive_discretization_3D.py:28-32`)."""
    return np.array([v for lim in _bounds(node) for v in lim], dtype=np.float64)


def _node_bounds(nodes) -> np.ndarray:
    return np.array([[v for lim in _bounds(n) for v in lim] for n in nodes], dtype=np.float64)


def _cheby_weights_leaf_order(bounds: np.ndarray, p: int) -> np.ndarray:
    """(n, p^d) tensor Clenshaw-Curtis weights in the boundary-first leaf ordering
    (`_adaptive_discretization_2D.py:380-407`, `_adaptive_discretization_3D.py:430-463`)."""
    b = np.asarray(bounds, dtype=np.float64)
    d = b.shape[1] // 2
    per_axis = [np.stack([chebyshev_weights(p, row[2 * a : 2 * a + 2]) for row in b]) for a in range(d)]
    w = per_axis[0]
    for a in range(1, d):
        w = (w[:, :, None] * per_axis[a][:, None, :]).reshape(b.shape[0], -1)
    r = rearrange_indices_ext_int_2D(p) if d == 2 else rearrange_indices_ext_int_3D(p)
    return w[:, r]


def get_squared_l2_norm_single_panel(f_evals, bounds, p: int) -> float:
    """Squared L2 norm over one 2D leaf from its Chebyshev samples (`_adaptive_discretization_2D.py:380-407`)."""
    return float(np.sum(_cheby_weights_leaf_order(np.asarray(bounds)[None], p)[0] * np.asarray(f_evals) ** 2))


def get_squared_l2_norm_single_voxel(f_evals, bounds, p: int) -> float:
    """Squared L2 norm over one 3D leaf (`_adaptive_discretization_3D.py:430-463`)."""
    return float(np.sum(_cheby_weights_leaf_order(np.asarray(bounds)[None], p)[0] * np.asarray(f_evals) ** 2))


def _position(root, node):
    """Integer position of ``node`` among the 2^depth boxes per axis at its depth."""
    k = 1 << node.depth
    return tuple(
        int(round((lo - rlo) / (rhi - rlo) * k)) for (lo, _), (rlo, rhi) in zip(_bounds(node), _bounds(root))
    )


def _ensure_box(root, depth: int, pos, q: int):
    """Make sure the box at (``depth``, ``pos``) exists; returns the node that had to be split for
    it, or None (`find_or_add_child`, `_adaptive_discretization_3D.py:300-395`)."""
    offs = _OFFSETS_2D if _is_2D(root) else _OFFSETS_3D
    cur = root
    for lvl in range(depth):
        if not cur.children:
            if lvl != depth - 1:
                raise ValueError("Requested volume is too large for the current node")
            _add_children(cur, root, q)
            return cur
        bits = tuple((c >> (depth - 1 - lvl)) & 1 for c in pos)
        cur = cur.children[offs.index(bits)]
    return None


def _generate(root, f_fn: Callable, tol: float, p: int, q: int, restrict_bool: bool, l2_norm: bool) -> None:
    two_d = _is_2D(root)
    d = 2 if two_d else 3
    to_pts = bounds_to_cheby_points_2D if two_d else bounds_to_cheby_points_3D
    split = quad_children_bounds if two_d else oct_children_bounds
    refine_op = precompute_L_4f1(p) if two_d else precompute_L_8f1(p)
    n_c = p**d

    def clouds(bounds):
        kids = split(bounds).reshape(-1, 2 * d)
        return to_pts(bounds, p), to_pts(kids, p).reshape(bounds.shape[0], -1, d), kids

    if l2_norm:
        _add_children(root, root, q)
        lb = _node_bounds(get_all_leaves(root))
        vals = np.asarray(f_fn(to_pts(lb, p)))
        global_nrm = float(np.sum(_cheby_weights_leaf_order(lb, p) * vals**2))
        tol = tol**2
    else:
        # the reference seeds the norm with max(f), not max|f| (`_adaptive_discretization_3D.py:66-69`)
        _, fine, _ = clouds(_node_bounds([root]))
        global_nrm = float(np.max(np.asarray(f_fn(fine.reshape(-1, d)))))

    queue: List = list(get_all_leaves(root))
    while queue:
        logging.debug("adaptive mesh: queue length %d", len(queue))
        qb = _node_bounds(queue)
        coarse, fine, kid_bounds = clouds(qb)
        f0 = np.asarray(f_fn(coarse), dtype=np.float64)
        f1 = np.asarray(f_fn(fine), dtype=np.float64)
        diff = f0 @ refine_op.T - f1
        if l2_norm:
            w = _cheby_weights_leaf_order(kid_bounds, p).reshape(len(queue), -1)
            ok = np.sum(w * diff**2, axis=1) / global_nrm < tol
        else:
            ref_max = np.max(np.abs(f1), axis=1)
            nrm = np.maximum(global_nrm, ref_max)
            ok = np.max(np.abs(diff), axis=1) / nrm < tol
            global_nrm = float(np.max(nrm))
        assert diff.shape[1] == (1 << d) * n_c
        nxt: List = []
        for node, good in zip(queue, ok):
            if good:
                continue
            _add_children(node, root, q)
            nxt.extend(node.children)
            if not restrict_bool:
                continue
            pending = [node]
            while pending:  # every same-size neighbour of a split box must exist
                cur = pending.pop()
                pos, k = _position(root, cur), 1 << cur.depth
                for ax in range(d):
                    for step in (-1, 1):
                        nb = list(pos)
                        nb[ax] += step
                        if 0 <= nb[ax] < k:
                            made = _ensure_box(root, cur.depth, tuple(nb), q)
                            if made is not None:
                                pending.append(made)
                                nxt.extend(made.children)
        queue = nxt


def generate_adaptive_mesh_level_restriction_2D(root, f_fn, tol, p, q, restrict_bool=True, l2_norm=False) -> None:
    """Refine ``root`` in place (`_adaptive_discretization_2D.py:37-199`)."""
    _generate(root, f_fn, tol, p, q, restrict_bool, l2_norm)


def generate_adaptive_mesh_level_restriction_3D(root, f_fn, tol, p, q, restrict_bool=True, l2_norm=False) -> None:
    """Refine ``root`` in place (`_adaptive_discretization_3D.py:35-205`)."""
    _generate(root, f_fn, tol, p, q, restrict_bool, l2_norm)


generate_adaptive_mesh_level_restriction = generate_adaptive_mesh_level_restriction_3D
