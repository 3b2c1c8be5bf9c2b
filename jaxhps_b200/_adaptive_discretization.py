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


class _DeviceCheck:
    """The per-round refinement criterion on the GPU (``hps_refine_check``): the interpolation of the whole queue to
    the children's clouds is one DMMA GEMM, the comparison one reduction per box; the samples of ``f`` travel host ->
    device (``f`` is a host callable, as in the reference), three numbers per box come back."""

    def __init__(self, refine_op: np.ndarray, device):
        import torch

        from . import _lib

        self.torch, self._lib = torch, _lib
        self.dev = _lib.require_cuda(device)
        self.lib = _lib.load()
        self.LT = _lib.to_device(np.ascontiguousarray(refine_op.T), self.dev)

    def __call__(self, f0: np.ndarray, f1: np.ndarray, w):
        import ctypes

        torch, _lib = self.torch, self._lib
        n, n_c = f0.shape
        n_f = f1.shape[1]
        with torch.cuda.device(self.dev):
            d0, d1 = _lib.to_device(f0, self.dev), _lib.to_device(f1, self.dev)
            dw = None if w is None else _lib.to_device(w, self.dev)
            out = torch.empty((3, n), dtype=torch.float64, device=self.dev)
            need = ctypes.c_size_t()
            _lib.check(self.lib.hps_refine_check_workspace(n, n_f, ctypes.byref(need)), "workspace query")
            ws = _lib.workspace(need.value, self.dev)
            rc = self.lib.hps_refine_check(_lib.stream_ptr(), n, n_c, n_f, _lib.ptr(d0), _lib.ptr(d1), _lib.ptr(self.LT),
                                           _lib.ptr(dw), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), _lib.ptr(ws),
                                           ws.numel())
            _lib.check(rc, "hps_refine_check")
            res = out.cpu().numpy()
        return res[0], res[1], res[2]


def _generate(root, f_fn: Callable, tol: float, p: int, q: int, restrict_bool: bool, l2_norm: bool, device=None) -> None:
    """``device``: a CUDA device runs the per-round criterion through ``hps_refine_check``; ``None`` keeps it in NumPy
    (the two agree to rounding: `tests/test_gpu_adaptive.py`)."""
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

    check = _DeviceCheck(refine_op, device) if device is not None else None
    queue: List = list(get_all_leaves(root))
    while queue:
        logging.debug("adaptive mesh: queue length %d", len(queue))
        qb = _node_bounds(queue)
        coarse, fine, kid_bounds = clouds(qb)
        f0 = np.asarray(f_fn(coarse), dtype=np.float64)
        f1 = np.asarray(f_fn(fine), dtype=np.float64)
        assert f1.shape[1] == (1 << d) * n_c
        w = _cheby_weights_leaf_order(kid_bounds, p).reshape(len(queue), -1) if l2_norm else None
        if check is not None:
            err_inf, err_l2, ref_max = check(np.ascontiguousarray(f0), np.ascontiguousarray(f1), w)
        else:
            diff = f0 @ refine_op.T - f1
            err_inf = np.max(np.abs(diff), axis=1)
            err_l2 = np.sum(w * diff**2, axis=1) if l2_norm else None
            ref_max = np.max(np.abs(f1), axis=1)
        if l2_norm:
            ok = err_l2 / global_nrm < tol
        else:
            nrm = np.maximum(global_nrm, ref_max)
            ok = err_inf / nrm < tol
            global_nrm = float(np.max(nrm))
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


def generate_adaptive_mesh_level_restriction_2D(root, f_fn, tol, p, q, restrict_bool=True, l2_norm=False, device=None) -> None:
    """Refine ``root`` in place (`_adaptive_discretization_2D.py:37-199`); ``device``: run the refinement check on a GPU."""
    _generate(root, f_fn, tol, p, q, restrict_bool, l2_norm, device)


def generate_adaptive_mesh_level_restriction_3D(root, f_fn, tol, p, q, restrict_bool=True, l2_norm=False, device=None) -> None:
    """Refine ``root`` in place (`_adaptive_discretization_3D.py:35-205`); ``device``: run the refinement check on a GPU."""
    _generate(root, f_fn, tol, p, q, restrict_bool, l2_norm, device)


generate_adaptive_mesh_level_restriction = generate_adaptive_mesh_level_restriction_3D
