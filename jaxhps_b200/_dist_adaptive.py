"""Subtree-sharded build + solve of an ADAPTIVE tree over several GPUs of one box (one process per GPU).

Same scheme as ``_dist.py`` (the parallel form of the reference's serial subtree loop,
`src/jaxhps/_subtree_recomp.py:310-390`), for non-uniform trees (BASELINE config 5):

* rank ``r`` of ``world`` owns the root's children ``[r*k, (r+1)*k)``, ``k = n_children / world`` — a
  contiguous leaf range (depth-first leaf order); it runs the leaf solves and every merge inside its
  subtrees with no communication and keeps their ``Y, v, S, g_tilde`` resident.  Adaptive trees are
  unbalanced: ``AdaptiveShardPlan.leaves_per_rank`` says by how much;
* up: every owner coarsens its subtree roots' ``(T, h)`` for the root interfaces and broadcasts them
  (NCCL over NVLink) — the root merge needs blocks of all children;
* the root merge is column-sharded: every rank assembles the root's interface system; below 8192 unknowns each
  rank factors it on its own (replicated LU), above the ranks factor it together (``_dist.distributed_lu_solve``:
  block columns dealt round-robin, NCCL broadcast per block column, look-ahead); each rank then solves only its
  share of the exterior columns of ``S``;
* down: ``g_int = sum_r S[:, cols_r] g[cols_r] + g_tilde`` is one all-reduce of an ``n_int`` vector; every
  rank then scatters the root data to its own children and runs its subtrees' down passes.

The arithmetic is delegated to an ``ops`` object: :class:`CudaAdaptiveOps` (the product) or the
oracle-backed test double of ``tests/test_dist_adaptive_gloo.py``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List

import numpy as np
import torch
import torch.distributed as dist

from ._adaptive_plan import TreePlan, get_plan
from ._pdeproblem import _get_PDEProblem_chunk
from ._tree import get_all_leaves

_PRIVATE_KEYS = ("_tree_plan", "_adaptive_state", "_adaptive_leaf", "_device_constants")


class AdaptiveShardPlan:
    """Which children of the root (and which leaves) each rank owns."""

    def __init__(self, root, rank: int, world: int):
        n_child = len(root.children)
        if n_child == 0:
            raise ValueError("the root has no children: nothing to shard")
        if world < 1 or n_child % world:
            raise ValueError(f"the number of ranks must divide the root's {n_child} children")
        if not 0 <= rank < world:
            raise ValueError("rank out of range")
        self.rank, self.world, self.n_child = rank, world, n_child
        self.per_rank = n_child // world
        self.children = list(range(rank * self.per_rank, (rank + 1) * self.per_rank))
        counts = [len(get_all_leaves(c)) for c in root.children]
        self.leaf_off = np.concatenate([[0], np.cumsum(counts)]).astype(int)
        self.leaves_per_rank = [int(sum(counts[r * self.per_rank : (r + 1) * self.per_rank])) for r in range(world)]
        self.leaf_slice = slice(int(self.leaf_off[self.children[0]]), int(self.leaf_off[self.children[-1] + 1]))

    def owner(self, child: int) -> int:
        return child // self.per_rank

    def columns(self, n_ext_panels: int, rank: int = None):
        """This rank's share [e0, e1) of the root's exterior panels (columns of S)."""
        r = self.rank if rank is None else rank
        return (n_ext_panels * r) // self.world, (n_ext_panels * (r + 1)) // self.world


class _SubDomain:
    """What the stage functions read from ``pde_problem.domain`` for one subtree."""

    def __init__(self, domain, node):
        self.p, self.q, self.root = domain.p, domain.q, node
        self.bool_2D, self.bool_uniform, self.L = domain.bool_2D, False, None
        self.n_leaves = len(get_all_leaves(node))


def subtree_problem(pde_problem, shard: AdaptiveShardPlan, child: int):
    """The part of ``pde_problem`` living in the subtree of root child ``child``."""
    s, e = int(shard.leaf_off[child]), int(shard.leaf_off[child + 1])
    sub = _get_PDEProblem_chunk(pde_problem, s, e)
    for key in _PRIVATE_KEYS:
        sub.__dict__.pop(key, None)
    sub.domain = _SubDomain(pde_problem.domain, pde_problem.domain.root.children[child])
    sub.reset()
    return sub


class CudaAdaptiveOps:
    """Device arithmetic of the sharded adaptive driver (CUDA kernels through the C ABI)."""

    def __init__(self, device):
        from . import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.dev = _lib.require_cuda(device)
        self._tables = None
        self._resident = {}  # host array id -> device copy (kept alive: raw pointers are handed to the kernels)

    def _dev(self, x):
        if isinstance(x, torch.Tensor):
            return x
        key = id(x)
        if key not in self._resident:
            self._resident[key] = (x, self._lib.to_device(x, self.dev))
        return self._resident[key][1]

    # -- plumbing
    def to_array(self, x):
        return self._lib.to_device(x, self.dev)

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.float64, device=self.dev)

    def broadcast(self, arr, src: int):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(arr, src=src)
        return arr

    def all_reduce(self, arr):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(arr)
        return arr

    def to_host(self, arr):
        return arr.cpu().numpy()

    def _root_tables(self, root_plan):
        if self._tables is None:
            flat = np.concatenate([root_plan.int_tbl.reshape(-1), root_plan.ext_tbl.reshape(-1), root_plan.down_tbl.reshape(-1)]
                                  + [ch.seg.reshape(-1) for ch in root_plan.children]).astype(np.int32)
            t = self._lib.to_device(flat, self.dev, dtype=torch.int32)
            off, at = {}, 0
            for key, arr in [("int", root_plan.int_tbl), ("ext", root_plan.ext_tbl), ("down", root_plan.down_tbl)] + \
                    [(f"seg{c}", ch.seg) for c, ch in enumerate(root_plan.children)]:
                off[key] = t.data_ptr() + 4 * at
                at += arr.size
            self._tables = (t, off)
        return self._tables[1]

    # -- arithmetic
    def build_subtree(self, sub):
        """Leaf solves + all merges of one subtree; returns its root's (T, h (n, n_src))."""
        from .adaptive import _local_solve_adaptive, _merge_adaptive

        dim = 2 if sub.domain.bool_2D else 3
        Y, T, v, h = _local_solve_adaptive(sub, dim, self.dev, self.dev)
        sub.__dict__["_adaptive_leaf"] = (Y, v)
        sub.Y, sub.v = Y, v
        if not sub.domain.root.children:  # the root's child is itself a leaf
            return T[0], (h[0] if h.ndim == 3 else h[0].unsqueeze(-1))
        _merge_adaptive(sub, T, h, self.dev, self.dev, return_T=True)
        d = sub.domain.root.data
        return d.T, (d.h if d.h.ndim == 2 else d.h.unsqueeze(-1))

    def compress(self, T, h, root_plan, c: int, L_refine, L_coarsen):
        ch = root_plan.children[c]
        if ch.identity:
            return T.contiguous(), h.contiguous()
        npp = root_plan.npp
        need = ctypes.c_size_t()
        self._lib.check(self.lib.hps_adaptive_compress_workspace(ch.n, ch.n_out // npp, npp, ctypes.byref(need)), "ws query")
        ws = self._lib.WORKSPACE.get(need.value, self.dev)
        T2, h2 = self.empty((ch.n_out, ch.n_out)), self.empty((ch.n_out, h.shape[1]))
        T, h = T.contiguous(), h.contiguous()
        rc = self.lib.hps_adaptive_compress(self._lib.stream_ptr(), npp, root_plan.group, h.shape[1], ch.n, T.data_ptr(),
                                            h.data_ptr(), ch.n_out // npp, self._root_tables(root_plan)[f"seg{c}"],
                                            self._dev(L_refine).data_ptr(), self._dev(L_coarsen).data_ptr(),
                                            T2.data_ptr(), h2.data_ptr(), ws.data_ptr(), ws.numel())
        self._lib.check(rc, "hps_adaptive_compress")
        return T2, h2

    #: interface size from which the root's system is factored by all ranks together (``hps_lu_dist_*``)
    DIST_LU_MIN_N = int(os.environ.get("HPS_DIST_LU_MIN_N", 8192))
    FORCE_DIST_LU = False  # tests: exercise the distributed path with a single rank

    def root_merge(self, Ts, hs, root_plan, e0: int, e1: int):
        """This rank's columns of the root ``S`` (n_int x (e1-e0)*npp) and the full ``g_tilde``."""
        npp, n_src = root_plan.npp, hs[0].shape[1]
        off = self._root_tables(root_plan)
        S = self.empty((root_plan.n_int, (e1 - e0) * npp))
        g = self.empty((root_plan.n_int, n_src))
        world = dist.get_world_size() if dist.is_initialized() else 1
        ptrs = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        lds = (ctypes.c_int * len(Ts))(*[t.shape[1] for t in Ts])
        if (world > 1 or self.FORCE_DIST_LU) and root_plan.n_int >= self.DIST_LU_MIN_N:
            from ._dist import distributed_lu_solve

            D = self.empty((root_plan.n_int, root_plan.n_int))
            rc = self.lib.hps_merge_adaptive_assemble(self._lib.stream_ptr(), npp, n_src, len(Ts), ptrs(Ts), ptrs(hs), lds,
                                                      root_plan.int_tbl.shape[0], off["int"], root_plan.ext_tbl.shape[0],
                                                      off["ext"], D.data_ptr(), S.data_ptr(), g.data_ptr(), e0, e1 - e0)
            self._lib.check(rc, "hps_merge_adaptive_assemble")
            distributed_lu_solve(self._lib, self.dev, D, [S, g], dist.get_rank() if world > 1 else 0, world)
            return S, g
        info = torch.zeros(1, dtype=torch.int32, device=self.dev)
        need = ctypes.c_size_t()
        self._lib.check(self.lib.hps_merge_adaptive_workspace(root_plan.n_int, root_plan.n_ext, 0, ctypes.byref(need)), "ws query")
        ws = self._lib.WORKSPACE.get(need.value, self.dev)
        def run():
            rc = self.lib.hps_merge_adaptive(self._lib.stream_ptr(), npp, n_src, len(Ts), ptrs(Ts), ptrs(hs), lds,
                                             root_plan.int_tbl.shape[0], off["int"], root_plan.ext_tbl.shape[0], off["ext"],
                                             S.data_ptr(), g.data_ptr(), None, None, 0, 0, None, e0, e1 - e0, ws.data_ptr(),
                                             ws.numel(), info.data_ptr())
            self._lib.check(rc, "hps_merge_adaptive (root columns)")
            self._lib.check_info(info, "root merge")

        self._lib.with_pivoting_fallback(run)  # the call only reads the children's operators: repeatable
        return S, g

    def matvec(self, S, x):
        """``S @ x`` with the library's bandwidth kernel (narrow x) / DMMA GEMM — no vendor BLAS on the path."""
        S, x = S.contiguous(), x.contiguous()
        M, K = S.shape
        N = x.shape[1]
        out = self.empty((M, N))
        rc = self.lib.hps_dgemm_strided_batched(self._lib.stream_ptr(), M, N, K, 1.0, S.data_ptr(), K, 0, x.data_ptr(), N, 0,
                                                0.0, out.data_ptr(), N, 0, 1)
        self._lib.check(rc, "hps_dgemm_strided_batched (root matvec)")
        return out

    def down_root(self, root_plan, g_ext, g_int, L_refine):
        """Boundary vectors of ALL children of the root from the reduced interface data."""
        n_src = g_ext.shape[1]
        outs = [self.empty((ch.n, n_src)) for ch in root_plan.children]
        ptrs = (ctypes.c_void_p * len(outs))(*[t.data_ptr() for t in outs])
        ws, g_ext = g_int.contiguous(), g_ext.contiguous()
        rc = self.lib.hps_down_adaptive(self._lib.stream_ptr(), root_plan.npp, n_src, root_plan.n_int, root_plan.n_ext, None,
                                        g_ext.data_ptr(), None, len(outs), ptrs, root_plan.down_tbl.shape[0],
                                        self._root_tables(root_plan)["down"], self._dev(L_refine).data_ptr(), ws.data_ptr())
        self._lib.check(rc, "hps_down_adaptive (root scatter)")
        return outs

    def down_subtree(self, sub, g):
        from .adaptive import _down_adaptive
        from .down_pass import leaf_apply

        if not sub.domain.root.children:
            Y, v = sub.__dict__["_adaptive_leaf"]
            return leaf_apply(Y, g.reshape(1, g.shape[0], -1), v.reshape(1, Y.shape[1], -1), self.dev)
        u = _down_adaptive(sub, g, self.dev, self.dev)
        return u if u.ndim == 3 else u.unsqueeze(-1)


def build_solver_sharded_adaptive(pde_problem, ops, rank: int = None, world: int = None) -> Dict:
    """Build on this rank's share of the tree + its columns of the root merge.  ``pde_problem`` is the full
    problem (host arrays) on every rank.  Returns the state ``solve_sharded_adaptive`` needs."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    dom = pde_problem.domain
    if dom.bool_uniform:
        raise ValueError("use _dist.build_solver_sharded for uniform trees")
    shard = AdaptiveShardPlan(dom.root, rank, world)
    plan: TreePlan = get_plan(pde_problem)
    root_plan = plan.by_id[id(dom.root)]
    Lr, Lc = (pde_problem.L_2f1, pde_problem.L_1f2) if dom.bool_2D else (pde_problem.L_4f1, pde_problem.L_1f4)
    subs, mine = {}, {}
    for c in shard.children:
        sub = subtree_problem(pde_problem, shard, c)
        T, h = ops.build_subtree(sub)
        subs[c] = sub
        mine[c] = ops.compress(T, h, root_plan, c, Lr, Lc)
        del T, h
    n_src = next(iter(mine.values()))[1].shape[1]
    Ts: List = []
    hs: List = []
    for c, ch in enumerate(root_plan.children):
        if shard.owner(c) == rank:
            T2, h2 = mine[c]
        else:
            T2, h2 = ops.empty((ch.n_out, ch.n_out)), ops.empty((ch.n_out, n_src))
        Ts.append(ops.broadcast(T2, shard.owner(c)))
        hs.append(ops.broadcast(h2, shard.owner(c)))
    e0, e1 = shard.columns(root_plan.ext_tbl.shape[0])
    S_r, g_tilde = ops.root_merge(Ts, hs, root_plan, e0, e1)
    del Ts, hs, mine
    return dict(shard=shard, root_plan=root_plan, subs=subs, S_r=S_r, g_tilde=g_tilde, cols=(e0, e1), L_refine=Lr)


def solve_sharded_adaptive(state: Dict, boundary_data, ops):
    """Down pass.  ``boundary_data``: list with one array per side / face of the root (or their
    concatenation), the same on every rank.  Returns ``(u_local, leaf_slice)``: the solution on this
    rank's leaves, ``(n_local_leaves, p^d, n_src)``."""
    shard, root_plan = state["shard"], state["root_plan"]
    if isinstance(boundary_data, (list, tuple)):
        boundary_data = np.concatenate([np.asarray(b) for b in boundary_data])
    g_ext = ops.to_array(np.asarray(boundary_data, dtype=np.float64).reshape(root_plan.n_ext, -1))
    npp = root_plan.npp
    e0, e1 = state["cols"]
    part = ops.matvec(state["S_r"], g_ext[e0 * npp : e1 * npp])
    if shard.rank == 0:
        part = part + state["g_tilde"]
    g_int = ops.all_reduce(part)
    kids = ops.down_root(root_plan, g_ext, g_int, state["L_refine"])
    out = [ops.down_subtree(state["subs"][c], kids[c]) for c in shard.children]
    u = torch.cat(out) if isinstance(out[0], torch.Tensor) else np.concatenate(out)
    return u, shard.leaf_slice
