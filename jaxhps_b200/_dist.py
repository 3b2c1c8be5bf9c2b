"""Subtree-sharded build + solve over several GPUs of one box (one process per GPU).

The reference is single-device; its only notion of splitting a tree is the serial "subtree
recomputation" loop (`src/jaxhps/_subtree_recomp.py:310-390`): contiguous leaf ranges are subtrees,
each is solved and merged independently up to its root, the subtree roots are merged on top, and
boundary data flows back down the same way.  This module runs that scheme in parallel:

* rank ``r`` of ``world`` (2, 4 or 8) owns the root octants ``[8r/world, 8(r+1)/world)`` — a
  contiguous leaf range because leaves are stored in depth-first sibling order
  (`_grid_creation_3D.py:20-21`); it runs the local solves and every merge below the root with no
  communication and keeps its ``Y, v, S, g_tilde`` resident;
* up: one all-gather of the subtree-root ``(T, h)`` (NCCL over NVLink);
* the root merge is column-sharded: every rank factors the root ``D`` (replicated LU) and solves
  only its ``24m/world`` columns of ``S``;
* down: each rank multiplies its column block of ``S`` with its slice of the boundary data, one
  all-reduce of the 12m-vector of interface values, then every rank continues down its own
  subtrees with no further communication.

The arithmetic is delegated to an ``ops`` object: :class:`CudaOps` (the product, CUDA kernels via
the C ABI) or a test double built on the CPU oracle for the ``gloo`` world-size-2 tests.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch
import torch.distributed as dist

from ._pdeproblem import _COEFF_NAMES, PDEProblem


class SubtreePlan:
    """Which part of a uniform octree of depth ``L`` a rank owns."""

    def __init__(self, L: int, rank: int, world: int):
        if world not in (1, 2, 4, 8):
            raise ValueError("subtree sharding supports 1, 2, 4 or 8 ranks (root octants per rank must be whole)")
        if L < 1:
            raise ValueError("need at least one level to shard by root octant")
        if not 0 <= rank < world:
            raise ValueError("rank out of range")
        self.L, self.rank, self.world = L, rank, world
        self.octants_per_rank = 8 // world
        self.first_octant = rank * self.octants_per_rank
        self.leaves_per_octant = 8 ** (L - 1)
        self.n_local_leaves = self.octants_per_rank * self.leaves_per_octant
        self.leaf_slice = slice(self.first_octant * self.leaves_per_octant,
                                (self.first_octant + self.octants_per_rank) * self.leaves_per_octant)

    def column_window(self, n_ext: int):
        """This rank's share of the root's exterior unknowns (columns of the root ``S``)."""
        per = n_ext // self.world
        if per * self.world != n_ext:
            raise ValueError("root boundary size must divide by the number of ranks")
        return self.rank * per, per


def local_problem(domain, plan: SubtreePlan, source, **coeffs) -> PDEProblem:
    """A ``PDEProblem`` carrying only this rank's leaves (``source``/coefficients already sliced to
    ``plan.leaf_slice``); constant operators are those of the full domain."""
    full_shape = domain.interior_points[..., 0].shape
    n_c = full_shape[1]
    for name, val in list(coeffs.items()) + [("source", source)]:
        if val is not None and tuple(val.shape[:2]) != (plan.n_local_leaves, n_c):
            raise ValueError(f"{name}: expected leading shape {(plan.n_local_leaves, n_c)}, got {tuple(val.shape)}")
    pb = PDEProblem.__new__(PDEProblem)
    pb.__dict__.update(_operator_template(domain).__dict__)  # constant operators are shared
    for k in _COEFF_NAMES:
        setattr(pb, k, coeffs.get(k))
    pb.source = source
    pb.reset()
    return pb


_TEMPLATES = {}


def _operator_template(domain) -> PDEProblem:
    """Constant operators (D1, P, Q, ...) for ``domain``, built once per process."""
    key = id(domain)
    if key not in _TEMPLATES:
        shp = domain.interior_points[..., 0].shape
        zero = np.zeros(shp)
        _TEMPLATES[key] = PDEProblem(domain, source=zero, D_xx_coefficients=zero)
    return _TEMPLATES[key]


class CudaOps:
    """Device arithmetic of the sharded driver: thin calls into the stage shims / C ABI."""

    def __init__(self, device):
        from . import _lib

        self._lib = _lib
        self.dev = _lib.require_cuda(device)

    def tensor(self, x):
        return self._lib.to_device(x, self.dev)

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.float64, device=self.dev)

    def local_solve(self, pb):
        from .local_solve import local_solve_stage_uniform_3D_DtN

        return local_solve_stage_uniform_3D_DtN(pb, device=self.dev, host_device=self.dev)

    def merge_subtrees(self, T, h, levels: int, n_roots: int):
        from .merge import merge_subtrees_3D_DtN

        return merge_subtrees_3D_DtN(T, h, levels, n_roots, device=self.dev)

    def root_columns(self, T8, h8, col0: int, ncols: int):
        from .merge import merge_root_columns_3D_DtN

        return merge_root_columns_3D_DtN(T8, h8, col0, ncols, device=self.dev)

    def matvec(self, S_cols, g_slice):
        return S_cols @ g_slice

    def root_scatter(self, g_ext, g_int):
        """(24m, n_src), (12m, n_src) -> (8, 6m, n_src)."""
        lib = self._lib.load()
        m = g_int.shape[0] // 12
        n_src = g_int.shape[-1]
        out = self.empty((8, 6 * m, n_src))
        rc = lib.hps_down_oct_scatter(self._lib.stream_ptr(), 1, m, n_src, g_ext.data_ptr(), g_int.data_ptr(),
                                      out.data_ptr())
        self._lib.check(rc, "hps_down_oct_scatter")
        return out

    def down_local(self, g_roots, S_lst, g_lst, Y, v):
        from .down_pass import down_levels, leaf_apply

        n_src = g_roots.shape[-1]
        g_leaf = down_levels(g_roots.contiguous(), S_lst, [g.reshape(g.shape[0], g.shape[1], n_src) for g in g_lst],
                             3, self.dev)
        return leaf_apply(Y, g_leaf, v.reshape(v.shape[0], v.shape[1], n_src), self.dev)


class ShardedState:
    """What one rank keeps after the sharded build."""

    def __init__(self):
        self.Y = self.v = None
        self.S_lst: List = []
        self.g_tilde_lst: List = []
        self.S_root_cols = None  # (12m, 24m/world)
        self.g_tilde_root = None  # (12m[, n_src])
        self.col0 = self.ncols = 0


def _group_ok(plan: SubtreePlan) -> bool:
    return plan.world > 1 and dist.is_available() and dist.is_initialized()


def build_solver_sharded(pde_problem: PDEProblem, plan: SubtreePlan, device=None, ops=None, group=None) -> ShardedState:
    """Local solves + merges on this rank's subtrees, all-gather of the subtree roots, column-sharded
    root merge.  ``pde_problem`` holds this rank's leaves only (see :func:`local_problem`)."""
    ops = ops or CudaOps(device)
    st = ShardedState()
    Y, T, v, h = ops.local_solve(pde_problem)
    st.Y, st.v = Y, v
    n_oct = plan.octants_per_rank
    if plan.L > 1:
        st.S_lst, st.g_tilde_lst, T_roots, h_roots = ops.merge_subtrees(T, h, plan.L - 1, n_oct)
    else:
        T_roots, h_roots = T, h
    # ---- up: gather the 8 subtree-root operators on every rank ----
    if plan.world > 1:
        if not _group_ok(plan):
            raise RuntimeError("torch.distributed must be initialised for world > 1")
        T8 = ops.empty((8,) + tuple(T_roots.shape[1:]))
        h8 = ops.empty((8,) + tuple(h_roots.shape[1:]))
        dist.all_gather_into_tensor(T8, T_roots.contiguous(), group=group)
        dist.all_gather_into_tensor(h8, h_roots.contiguous(), group=group)
    else:
        T8, h8 = T_roots, h_roots
    del T_roots, h_roots
    n_ext = 4 * T8.shape[-1]  # 24 m
    st.col0, st.ncols = plan.column_window(n_ext)
    st.S_root_cols, st.g_tilde_root = ops.root_columns(T8, h8, st.col0, st.ncols)
    return st


def solve_sharded(pde_problem: PDEProblem, st: ShardedState, plan: SubtreePlan, boundary_data, device=None, ops=None,
                  group=None):
    """Down pass: returns this rank's part of the solution, ``(n_local_leaves, p^3[, n_src])`` as a
    tensor on the compute device.  ``boundary_data`` is the full root boundary vector."""
    ops = ops or CudaOps(device)
    g = ops.tensor(boundary_data)
    single = g.ndim == 1
    g_ext = g.reshape(g.shape[0], -1)
    gt = st.g_tilde_root.reshape(st.g_tilde_root.shape[0], -1)
    # ---- root level: partial product on this rank's columns, all-reduce, add g~ ----
    part = ops.matvec(st.S_root_cols, g_ext[st.col0 : st.col0 + st.ncols].contiguous())
    if plan.world > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    g_int = part + gt
    kids = ops.root_scatter(g_ext.contiguous(), g_int.contiguous())
    mine = kids[plan.first_octant : plan.first_octant + plan.octants_per_rank]
    u = ops.down_local(mine, st.S_lst, st.g_tilde_lst, st.Y, st.v)
    return u[..., 0] if single else u
