"""Subtree-sharded build + solve over several GPUs of one box (one process per GPU).

The reference is single-device; its only notion of splitting a tree is the serial "subtree
recomputation" loop (`src/jaxhps/_subtree_recomp.py:310-390`): contiguous leaf ranges are subtrees,
each is solved and merged independently up to its root, the subtree roots are merged on top, and
boundary data flows back down the same way.  This module runs that scheme in parallel:

* rank ``r`` of ``world`` (2, 4 or 8) owns the root octants ``[8r/world, 8(r+1)/world)`` — a
  contiguous leaf range because leaves are stored in depth-first sibling order
  (`_grid_creation_3D.py:20-21`); it runs the local solves and every merge below the root with no
  communication and keeps its ``Y, v, S, g_tilde`` resident;
* up: one all-gather of the subtree roots' interface blocks (NCCL over NVLink);
* the root merge is column-sharded BY CHILD: the columns of the root ``S`` that belong to child X's
  exterior faces depend on ``T_X`` only, so only the children's interface blocks ``T_X[int,int]``
  (a quarter of ``T_X``) are all-gathered; every rank factors the root ``D`` (replicated LU) and
  solves the columns of its own children;
* down: each rank multiplies its column block of ``S`` with its slice of the boundary data, one
  all-reduce of the 12m-vector of interface values, then every rank continues down its own
  subtrees with no further communication.

The arithmetic is delegated to an ``ops`` object: :class:`CudaOps` (the product, CUDA kernels via
the C ABI) or a test double built on the CPU oracle for the ``gloo`` world-size-2 tests.
"""
from __future__ import annotations

import os
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from ._pdeproblem import _COEFF_NAMES, PDEProblem


# exterior faces of the root's children a..h (ascending) and, per parent face, the four children
# in panel order (reference `merge/_uniform_3D_DtN.py:238-380, 507-541`)
_CHILD_EXT_FACES = [(0, 2, 5), (1, 2, 5), (1, 3, 5), (0, 3, 5), (0, 2, 4), (1, 2, 4), (1, 3, 4), (0, 3, 4)]
_FACE_CHILDREN = [(4, 7, 3, 0), (5, 6, 2, 1), (4, 5, 1, 0), (7, 6, 2, 3), (4, 5, 6, 7), (0, 1, 2, 3)]


def child_column_index(first_child: int, n_children: int, m: int) -> np.ndarray:
    """Positions, inside the root's face-ordered boundary vector (24 panels of m), of the exterior
    unknowns of the given children — child-major, faces ascending: the column order of ``S_r``."""
    idx = []
    for c in range(first_child, first_child + n_children):
        for f in _CHILD_EXT_FACES[c]:
            panel = 4 * f + _FACE_CHILDREN[f].index(c)
            idx.append(np.arange(panel * m, (panel + 1) * m))
    return np.concatenate(idx)


#: first interface (slot 0..11 of the root's interfaces 9..20) touched by the root's children a..h: the rows of
#: -C above it are structurally zero in that child's columns (`hps_root_cols_structure`)
_FIRST_SLOT = [0, 0, 1, 2, 4, 4, 5, 6]


def balanced_panels(world: int):
    """Which exterior panels ``(child, i)`` (the i-th exterior face of a root child: m columns of the root S) every
    rank solves for.  A column that starts with s/12 zero rows costs (1 - s/12)^2 of a full one in the forward
    substitution, so children a, b are the expensive ones and h the cheapest: the 24 panels are dealt out greedily,
    heaviest first, 24 / world per rank, and sorted by first interface inside a rank (the order the structured
    solve needs).  With one rank this is the reference's region order a..h."""
    w = [(1.0 - s / 12.0) ** 2 for s in _FIRST_SLOT]
    panels = sorted(((c, i) for c in range(8) for i in range(3)), key=lambda p: (-w[p[0]], p))
    cap = 24 // world
    bins, load = [[] for _ in range(world)], [0.0] * world
    for p in panels:
        r = min((q for q in range(world) if len(bins[q]) < cap), key=lambda q: (load[q], q))
        bins[r].append(p)
        load[r] += w[p[0]]
    return [sorted(b, key=lambda p: (_FIRST_SLOT[p[0]], p)) for b in bins]


def panel_column_index(panels, m: int) -> np.ndarray:
    """Positions of the given panels' unknowns inside the root's face-ordered boundary vector."""
    idx = []
    for c, i in panels:
        f = _CHILD_EXT_FACES[c][i]
        panel = 4 * f + _FACE_CHILDREN[f].index(c)
        idx.append(np.arange(panel * m, (panel + 1) * m))
    return np.concatenate(idx)


def exchange_panels(ops, Cblk, plan: "SubtreePlan", assign, m: int, group=None):
    """``Cpan`` (n_panels, 3m, m) of this rank's panels, in ``assign[rank]`` order: its own children's panels are
    sliced out of ``Cblk`` (n_local, 3m, 3m), the others arrive through one all-to-all (each rank sends most of its
    three-per-child panels away; 61 MB per panel at L=3, under 1 GB at L=4 — NVLink-trivial next to the solve)."""
    world, rank, per = plan.world, plan.rank, plan.octants_per_rank
    first = plan.first_octant
    mine = assign[rank]

    def cut(c, i):
        return Cblk[c - first][:, i * m : (i + 1) * m]

    if world == 1:
        return torch.stack([cut(c, i) for c, i in mine]).contiguous()
    send_parts, send_counts = [], []
    for q in range(world):
        ps = [p for p in assign[q] if p[0] // per == rank]
        send_counts.append(len(ps))
        send_parts += [cut(c, i) for c, i in ps]
    send = torch.stack(send_parts).contiguous() if send_parts else ops.empty((0, 3 * m, m))
    arrival = [p for q in range(world) for p in mine if p[0] // per == q]  # order in which the panels arrive
    recv_counts = [sum(1 for p in mine if p[0] // per == q) for q in range(world)]
    recv = ops.empty((len(mine), 3 * m, m))
    dist.all_to_all_single(recv, send, recv_counts, send_counts, group=group)
    order = [arrival.index(p) for p in mine]
    if order == list(range(len(mine))):
        return recv
    return recv[torch.as_tensor(order, device=recv.device)].contiguous()


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


class P2PComm:
    """The library's symmetric-segment communicator (``hps_comm_*``): one per (process, device).  The 64-byte CUDA
    IPC handles of the ranks' segments are exchanged once per (re)allocation with ``torch.distributed``; after
    that every exchange of the distributed root factorisation is a kernel storing into the peers' HBM."""

    _instances = {}

    def __init__(self, _lib, dev, rank: int, world: int, group=None):
        import ctypes

        self._lib, self.dev, self.rank, self.world, self.group = _lib, dev, rank, world, group
        self.lib = _lib.load()
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.lib.hps_comm_create(rank, world, ctypes.byref(self.handle)), "hps_comm_create")

    @classmethod
    def get(cls, _lib, dev, rank: int, world: int, group=None):
        key = (str(dev), rank, world, id(group))
        if key not in cls._instances:
            cls._instances[key] = cls(_lib, dev, rank, world, group)
        return cls._instances[key]

    def ensure(self, nbytes: int) -> None:
        """Collective: make every rank's segment at least ``nbytes`` large and mapped by all peers."""
        import ctypes

        lib, _lib = self.lib, self._lib
        with torch.cuda.device(self.dev):
            changed = ctypes.c_int(0)
            rc = lib.hps_comm_reserve(self.handle, nbytes, ctypes.byref(changed))
            if rc != 0:  # a mapped segment has to grow: unmap everywhere first
                _lib.check(lib.hps_comm_detach(self.handle), "hps_comm_detach")
                if self.world > 1:
                    torch.cuda.synchronize(self.dev)
                    dist.barrier(group=self.group)
                _lib.check(lib.hps_comm_reserve(self.handle, nbytes, ctypes.byref(changed)), "hps_comm_reserve")
            if not changed.value:
                return
            mine = (ctypes.c_ubyte * 64)()
            _lib.check(lib.hps_comm_export(self.handle, mine), "hps_comm_export")
            if self.world > 1:
                h = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
                allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.dev)
                dist.all_gather_into_tensor(allh, h, group=self.group)
                raw = bytes(allh.cpu().tolist())
            else:
                raw = bytes(mine)
            buf = (ctypes.c_ubyte * len(raw)).from_buffer_copy(raw)
            _lib.check(lib.hps_comm_attach(self.handle, buf), "hps_comm_attach")
            if self.world > 1:
                dist.barrier(group=self.group)

    def matrix_ptr(self, n: int) -> int:
        import ctypes

        out = ctypes.c_void_p()
        self._lib.check(self.lib.hps_lu_dist_matrix_ptr(self.handle, n, ctypes.byref(out)), "hps_lu_dist_matrix_ptr")
        return out.value

    def lu_segment_bytes(self, n: int) -> int:
        import ctypes

        need = ctypes.c_size_t()
        self._lib.check(self.lib.hps_lu_dist_segment_bytes(n, ctypes.byref(need)), "hps_lu_dist_segment_bytes")
        return need.value


#: record CUDA events at the stage boundaries of build_solver_sharded (``ShardedState.timing``, ms; one extra sync)
STAGE_TIMING = os.environ.get("HPS_DIST_TIMING", "0") == "1"

#: the distributed factorisation runs from C over the library's P2P communicator (0: the step-wise
#: NCCL-broadcast driver below, kept as the fallback when CUDA IPC is unavailable)
USE_P2P = os.environ.get("HPS_DIST_P2P", "1") != "0"


class StructureInvalid(Exception):
    """The factorisation interchanged rows: a solve that skipped declared zero rows must be repeated without."""


def root_cols_structure(_lib, first_child: int, n_local: int, m: int):
    """``(n_seg, seg_cols, seg_first_row)`` of the root's ``-C_r`` for children ``first_child ..`` (child-major
    columns): a child's exterior columns are zero above that child's first interface (``hps_root_cols_structure``)."""
    import ctypes

    n_seg, seg_cols = ctypes.c_int(), ctypes.c_int()
    first = (ctypes.c_int * (3 * n_local))()
    _lib.check(_lib.load().hps_root_cols_structure(first_child, n_local, m, ctypes.byref(n_seg), ctypes.byref(seg_cols), first),
               "hps_root_cols_structure")
    return n_seg.value, seg_cols.value, first


def p2p_lu_solve(_lib, dev, comm: "P2PComm", n: int, rhs, group=None, structure=None) -> None:
    """``rhs[k] := D^-1 rhs[k]`` with D already assembled at ``comm.matrix_ptr(n)`` on every rank
    (``hps_lu_dist_run``: one C call enqueues the whole factorisation, the P2P exchanges and the solves).
    ``structure``: leading-zero description of ``rhs[0]`` (:func:`root_cols_structure`); raises
    :class:`StructureInvalid` when the factorisation moved rows (the caller re-assembles and solves without it)."""
    import ctypes

    lib = _lib.load()
    need = ctypes.c_size_t()
    _lib.check(lib.hps_lu_solve_workspace(1, n, ctypes.byref(need)), "workspace query")
    ws = _lib.workspace(need.value, dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    k = len(rhs)
    ptrs = (ctypes.c_void_p * k)(*[r.data_ptr() for r in rhs])
    lds = (ctypes.c_int64 * k)(*[r.shape[1] for r in rhs])
    ncs = (ctypes.c_int * k)(*[r.shape[1] for r in rhs])
    if structure is not None:
        n_seg, seg_cols, first = structure
        _lib.check(lib.hps_lu_dist_run_structured(comm.handle, _lib.stream_ptr(), n, k, ptrs, lds, ncs, n_seg, seg_cols, first,
                                                  ws.data_ptr(), ws.numel(), info.data_ptr()), "hps_lu_dist_run_structured")
    else:
        _lib.check(lib.hps_lu_dist_run(comm.handle, _lib.stream_ptr(), n, k, ptrs, lds, ncs, ws.data_ptr(), ws.numel(),
                                       info.data_ptr()), "hps_lu_dist_run")
    # zero pivots (> 0) are seen by the owner of the block column only, and so are the assumption codes (< 0: -1 rows
    # moved under a structured solve, -2 a speculative block column needed pivoting): reduce both signs separately
    codes = torch.stack([info[0].clamp(min=0), (-info[0]).clamp(min=0)])
    if comm.world > 1:
        dist.all_reduce(codes, op=dist.ReduceOp.MAX, group=group)
    pos, neg = int(codes[0]), int(codes[1])
    if pos:
        raise np.linalg.LinAlgError(f"distributed factorisation: exact zero pivot at column {pos}")
    if neg >= 2:
        raise _lib.AssumptionViolated("distributed factorisation")
    if neg == 1:
        raise StructureInvalid()


def distributed_lu_solve(_lib, dev, D, rhs, rank: int, world: int, group=None) -> None:
    """In-place ``rhs[k] := D^-1 rhs[k]`` with the factorisation of ``D`` (n x n, replicated on every rank)
    distributed over the ranks: block column b (128 wide) is factored by rank ``b % world``, the packed column is
    broadcast with NCCL and every rank applies it to the block columns it owns; a high-priority side stream keeps
    the panel chain (update of the NEXT block column, its factorisation, the broadcast) one block ahead of the
    main stream's rank-128 updates (look-ahead).  Afterwards every rank holds the complete factors and runs the
    substitutions on its own right-hand sides (each rank may pass different ones)."""
    import ctypes

    lib = _lib.load()
    n = D.shape[0]
    if USE_P2P:
        comm = P2PComm.get(_lib, dev, rank, world, group)
        comm.ensure(comm.lu_segment_bytes(n))
        _copy_into_segment(comm.matrix_ptr(n), D)
        # a general matrix (the adaptive root's interface system): the right-hand sides are overwritten in place, so a
        # failed speculation could not be repeated — factor with full partial pivoting from the start
        with _lib.speculation(False):
            return p2p_lu_solve(_lib, dev, comm, n, rhs, group)
    NB = 128
    nblk = (n + NB - 1) // NB
    need = ctypes.c_size_t()
    _lib.check(lib.hps_lu_solve_workspace(1, n, ctypes.byref(need)), "workspace query")
    ws = _lib.WORKSPACE.get(need.value, dev)
    cnt = ctypes.c_size_t()
    _lib.check(lib.hps_lu_dist_buffer_doubles(n, ctypes.byref(cnt)), "buffer query")
    bufs = [torch.empty((cnt.value,), dtype=torch.float64, device=dev) for _ in range(2)]
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    main = torch.cuda.current_stream()
    side = _side_stream(dev)
    ev_panel = [torch.cuda.Event() for _ in range(2)]
    ev_upd = [torch.cuda.Event() for _ in range(2)]
    A, wsp, wsn = D.data_ptr(), ws.data_ptr(), ws.numel()

    def owner_of(b):
        return b % world

    def bcast(buf, owner):
        if world > 1:
            dist.broadcast(buf, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)

    side.wait_stream(main)  # D and the right-hand sides are ready
    with torch.cuda.stream(side):
        if rank == owner_of(0):
            _lib.check(lib.hps_lu_dist_factor_pack(side.cuda_stream, n, A, n, 0, wsp, wsn, info.data_ptr(),
                                                   bufs[0].data_ptr()), "hps_lu_dist_factor_pack")
        bcast(bufs[0], owner_of(0))
        if rank != owner_of(0):
            _lib.check(lib.hps_lu_dist_unpack(side.cuda_stream, n, A, n, 0, wsp, wsn, bufs[0].data_ptr()), "hps_lu_dist_unpack")
        ev_panel[0].record(side)
    for b in range(nblk):
        nb1 = b + 1
        i_own_next = nb1 < nblk and rank == owner_of(nb1)
        if nb1 < nblk:
            with torch.cuda.stream(side):
                buf = bufs[nb1 & 1]
                if i_own_next:
                    if b >= 1:
                        side.wait_event(ev_upd[(b - 1) & 1])  # column nb1 has received blocks < b on the main stream
                    _lib.check(lib.hps_lu_dist_update(side.cuda_stream, n, A, n, b, nb1, 1, 1, 0, wsp, wsn),
                               "hps_lu_dist_update (look-ahead)")
                    _lib.check(lib.hps_lu_dist_factor_pack(side.cuda_stream, n, A, n, nb1, wsp, wsn, info.data_ptr(),
                                                           buf.data_ptr()), "hps_lu_dist_factor_pack")
                bcast(buf, owner_of(nb1))
                if not i_own_next:
                    _lib.check(lib.hps_lu_dist_unpack(side.cuda_stream, n, A, n, nb1, wsp, wsn, buf.data_ptr()),
                               "hps_lu_dist_unpack")
                ev_panel[nb1 & 1].record(side)
        # main stream: block b applied to this rank's remaining block columns (+ the left interchanges)
        main.wait_event(ev_panel[b & 1])
        start = nb1 + 1 if i_own_next else nb1
        first = start + ((rank - start) % world)
        n_own = 0 if first >= nblk else (nblk - 1 - first) // world + 1
        _lib.check(lib.hps_lu_dist_update(main.cuda_stream, n, A, n, b, first, n_own, world, 1, wsp, wsn),
                   "hps_lu_dist_update")
        ev_upd[b & 1].record(main)
    main.wait_stream(side)
    if world > 1:
        dist.all_reduce(info, op=dist.ReduceOp.MAX, group=group)
    _lib.check_info(info, "distributed factorisation")
    k = len(rhs)
    ptrs = (ctypes.c_void_p * k)(*[r.data_ptr() for r in rhs])
    lds = (ctypes.c_int64 * k)(*[r.shape[1] for r in rhs])
    ncs = (ctypes.c_int * k)(*[r.shape[1] for r in rhs])
    _lib.check(lib.hps_lu_dist_solve(main.cuda_stream, n, D.data_ptr(), n, k, ptrs, lds, ncs, ws.data_ptr(), ws.numel()),
               "hps_lu_dist_solve")


def _copy_into_segment(dst_ptr: int, D: torch.Tensor) -> None:
    """Device-to-device copy of a contiguous tensor to a raw device address, on the current stream."""
    from . import _lib

    D = D.contiguous()
    _lib.check(_lib.load().hps_memcpy_d2d(_lib.stream_ptr(), dst_ptr, D.data_ptr(), D.numel() * D.element_size()),
               "hps_memcpy_d2d")


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

    def root_pack(self, T_roots, h_roots, first_child: int):
        """(Dblk, Cblk, hblk) of this rank's subtree roots (``hps_root_pack_oct``)."""
        lib = self._lib.load()
        n_local, n6, _ = T_roots.shape
        m = n6 // 6
        h3 = h_roots.reshape(n_local, n6, -1).contiguous()
        n_src = h3.shape[-1]
        Dblk = self.empty((n_local, 3 * m, 3 * m))
        Cblk = self.empty((n_local, 3 * m, 3 * m))
        hblk = self.empty((n_local, 3 * m, n_src))
        rc = lib.hps_root_pack_oct(self._lib.stream_ptr(), n_local, first_child, m, n_src, T_roots.data_ptr(),
                                   h3.data_ptr(), Dblk.data_ptr(), Cblk.data_ptr(), hblk.data_ptr())
        self._lib.check(rc, "hps_root_pack_oct")
        return Dblk, Cblk, hblk

    #: distribute the factorisation of the root D over the ranks when it is at least this large
    DIST_LU_MIN_N = int(os.environ.get("HPS_DIST_LU_MIN_N", "8192"))
    #: tests: run the step-wise (distributed) factorisation even on a single rank
    FORCE_DIST_LU = False
    #: skip the structurally-zero leading rows of -C_r in the root's forward substitution (HPS_MERGE_STRUCT=0: off)
    STRUCTURED = os.environ.get("HPS_MERGE_STRUCT", "1") != "0"

    def root_solve(self, Dblk_all, hblk_all, Cpan, panels, rank: int = 0, world: int = 1, group=None,
                   root_mode: str = "S"):
        """This rank's columns of the root S (one block of m columns per panel in ``panels`` = [(child, i), ...],
        sorted by first interface) and the full g~.  ``Cpan`` (n_panels, 3m, m): see :func:`exchange_panels`.
        Single rank or small root: ``hps_root_solve_panels`` (replicated LU).  Otherwise the LU of D is distributed
        by block columns over the library's P2P communicator (``hps_lu_dist_run``; NCCL-broadcast fallback
        ``hps_lu_dist_*``).  ``root_mode="factored"``: S is not formed — returns ``-C_r`` in its place and leaves the
        factors of D in the communicator's segment for :meth:`root_apply`."""
        import ctypes

        lib = self._lib.load()
        n_pan, n3, m = Cpan.shape
        n_src = hblk_all.shape[-1]
        pc = (ctypes.c_int * n_pan)(*[c for c, _ in panels])
        if root_mode == "factored":
            if not USE_P2P:
                raise ValueError("root_mode='factored' keeps the factors in the P2P segment (HPS_DIST_P2P=0 disables it)")
            return self._root_solve_distributed(Dblk_all, hblk_all, Cpan, pc, rank, world, group, True)
        if (world > 1 or self.FORCE_DIST_LU) and 12 * m >= self.DIST_LU_MIN_N:
            return self._root_solve_distributed(Dblk_all, hblk_all, Cpan, pc, rank, world, group)
        S_r = self.empty((12 * m, n_pan * m))
        g = self.empty((12 * m, n_src))
        info = torch.zeros(1, dtype=torch.int32, device=self.dev)
        need = ctypes.c_size_t()
        self._lib.check(lib.hps_root_solve_oct_workspace(m, ctypes.byref(need)), "workspace query")
        ws = self._lib.WORKSPACE.get(need.value, self.dev)

        def run():
            rc = lib.hps_root_solve_panels(self._lib.stream_ptr(), m, n_src, n_pan, pc, Dblk_all.data_ptr(),
                                           hblk_all.data_ptr(), Cpan.data_ptr(), S_r.data_ptr(), g.data_ptr(), ws.data_ptr(),
                                           ws.numel(), info.data_ptr())
            self._lib.check(rc, "hps_root_solve_panels")
            self._lib.check_info(info, "root merge")

        self._lib.with_pivoting_fallback(run)
        return S_r, g

    def root_apply(self, x, rank: int, world: int, group=None):
        """In place ``x := D^-1 x`` with the factors of the last factored root build (``hps_lu_dist_apply``)."""
        import ctypes

        lib, _lib = self._lib.load(), self._lib
        comm = P2PComm.get(_lib, self.dev, rank, world, group)
        n = x.shape[0]
        need = ctypes.c_size_t()
        _lib.check(lib.hps_lu_solve_workspace(1, n, ctypes.byref(need)), "workspace query")
        ws = _lib.workspace(need.value, self.dev)
        ptrs = (ctypes.c_void_p * 1)(x.data_ptr())
        lds = (ctypes.c_int64 * 1)(x.shape[1])
        ncs = (ctypes.c_int * 1)(x.shape[1])
        _lib.check(lib.hps_lu_dist_apply(comm.handle, _lib.stream_ptr(), n, 1, ptrs, lds, ncs, ws.data_ptr(), ws.numel()),
                   "hps_lu_dist_apply")
        return x

    def _root_solve_distributed(self, Dblk_all, hblk_all, Cpan, pc, rank, world, group, factored=False):
        """Distributed LU of the root D: block column b is factored by rank b % world, stored into the peers'
        segments, and applied by every rank to the block columns it owns; the rank's right-hand sides ride along."""
        import ctypes

        lib = self._lib.load()
        _lib = self._lib
        n_pan, n3, m = Cpan.shape
        n_src = hblk_all.shape[-1]
        n = 12 * m
        S_r = self.empty((n, n_pan * m))
        g = self.empty((n, n_src))

        def assemble(D_ptr):
            rc = lib.hps_root_assemble_panels(_lib.stream_ptr(), m, n_src, n_pan, pc, Dblk_all.data_ptr(), hblk_all.data_ptr(),
                                              Cpan.data_ptr(), D_ptr, S_r.data_ptr(), g.data_ptr())
            _lib.check(rc, "hps_root_assemble_panels")

        if USE_P2P:
            # D is assembled directly inside this rank's symmetric segment; the owners of the block columns
            # later store the factored columns into the same place on every peer
            comm = P2PComm.get(_lib, self.dev, rank, world, group)
            comm.ensure(comm.lu_segment_bytes(n))
            structure = None
            if self.STRUCTURED and not factored:  # leading zero rows of -C_r, panel by panel
                n_seg, seg_cols = ctypes.c_int(), ctypes.c_int()
                first = (ctypes.c_int * n_pan)()
                _lib.check(lib.hps_root_panels_structure(n_pan, pc, m, ctypes.byref(n_seg), ctypes.byref(seg_cols), first),
                           "hps_root_panels_structure")
                structure = (n_seg.value, seg_cols.value, first)

            def run():
                assemble(comm.matrix_ptr(n))
                # factored root: only g~ goes through the solve; S_r keeps -C_r for the solves
                try:
                    p2p_lu_solve(_lib, self.dev, comm, n, [g] if factored else [S_r, g], group, structure)
                except StructureInvalid:  # rows were interchanged: the zero-row shortcut does not apply to this matrix
                    assemble(comm.matrix_ptr(n))
                    p2p_lu_solve(_lib, self.dev, comm, n, [g] if factored else [S_r, g], group)

            _lib.with_pivoting_fallback(run)  # (info = -2: the speculative block columns did not apply either)
            return S_r, g
        D = self.empty((n, n))
        assemble(D.data_ptr())
        distributed_lu_solve(_lib, self.dev, D, [S_r, g], rank, world, group)
        return S_r, g

    def matvec(self, S_cols, g_slice):
        """``S_cols @ g_slice`` with the library's bandwidth kernel (narrow N) / DMMA GEMM."""
        lib = self._lib.load()
        M, K = S_cols.shape
        N = g_slice.shape[1]
        out = self.empty((M, N))
        rc = lib.hps_dgemm_strided_batched(self._lib.stream_ptr(), M, N, K, 1.0, S_cols.data_ptr(), K, 0,
                                           g_slice.data_ptr(), N, 0, 0.0, out.data_ptr(), N, 0, 1)
        self._lib.check(rc, "hps_dgemm_strided_batched (root matvec)")
        return out

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


_SIDE_STREAMS = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    """One high-priority stream per device for the look-ahead panel chain."""
    key = str(dev)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _SIDE_STREAMS[key]


class ShardedState:
    """What one rank keeps after the sharded build."""

    def __init__(self):
        self.Y = self.v = None
        self.S_lst: List = []
        self.g_tilde_lst: List = []
        self.S_root_cols = None  # (12m, m * 24 / world): this rank's exterior panels (balanced_panels), m columns each
        self.panels = None  # [(child, i)] behind those column blocks
        self.g_tilde_root = None  # (12m, n_src)
        self.col_index = None  # where those columns sit in the root's boundary vector
        self.timing = None  # stage -> ms when STAGE_TIMING is on
        self.root_mode = "S"  # "factored": S_root_cols holds -C_r and the factors of D stay in the P2P segment


def _group_ok(plan: SubtreePlan) -> bool:
    return plan.world > 1 and dist.is_available() and dist.is_initialized()


def build_solver_sharded(pde_problem: PDEProblem, plan: SubtreePlan, device=None, ops=None, group=None,
                         root_mode: str = "S") -> ShardedState:
    """Local solves + merges on this rank's subtrees, all-gather of the subtree roots, column-sharded
    root merge.  ``pde_problem`` holds this rank's leaves only (see :func:`local_problem`).

    ``root_mode="S"`` (default) forms the root ``S`` like the reference's ``S_lst[-1]``.  ``root_mode="factored"``
    (opt-in; CUDA ops only) keeps ``P D = L U`` instead and applies ``D^-1`` to ``-C g - h_int`` in every solve:
    the right-hand-side substitutions that form ``S`` — 28 of the 41.7 TFLOP of an L=3 build, 226 of 320 per rank
    at L=4 on 8 GPUs — disappear from the build at the price of one pass over the factors per solve.  The factors
    live in the process's P2P segment: the state is valid until the next sharded build on this process."""
    if root_mode not in ("S", "factored"):
        raise ValueError("root_mode must be 'S' or 'factored'")
    ops = ops or CudaOps(device)
    st = ShardedState()
    st.root_mode = root_mode
    marks = [] if STAGE_TIMING else None  # developer aid (HPS_DIST_TIMING=1 / bench.py): CUDA events at stage boundaries

    def mark(name):
        if marks is not None and torch.cuda.is_available():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    mark("start")
    Y, T, v, h = ops.local_solve(pde_problem)
    mark("local_solve")
    st.Y, st.v = Y, v
    n_oct = plan.octants_per_rank
    if plan.L > 1:
        st.S_lst, st.g_tilde_lst, T_roots, h_roots = ops.merge_subtrees(T, h, plan.L - 1, n_oct)
    else:
        T_roots, h_roots = T, h
    mark("subtree_merges")
    # ---- up: only the children's interface blocks travel (a quarter of each subtree-root T) ----
    multi = h_roots.ndim == 3
    Dblk, Cblk, hblk = ops.root_pack(T_roots, h_roots, plan.first_octant)
    m = T_roots.shape[-1] // 6
    del T_roots, h_roots
    # which exterior panels this rank solves for: balanced over the ranks (see balanced_panels); one all-to-all
    assign = balanced_panels(plan.world)
    panels = assign[plan.rank]
    if plan.world > 1 and not _group_ok(plan):
        raise RuntimeError("torch.distributed must be initialised for world > 1")
    mark("root_pack")
    Cpan = exchange_panels(ops, Cblk, plan, assign, m, group)
    del Cblk
    mark("panel_exchange")
    if plan.world > 1:
        if not _group_ok(plan):
            raise RuntimeError("torch.distributed must be initialised for world > 1")
        D_all = ops.empty((8,) + tuple(Dblk.shape[1:]))
        h_all = ops.empty((8,) + tuple(hblk.shape[1:]))
        dist.all_gather_into_tensor(D_all, Dblk.contiguous(), group=group)
        dist.all_gather_into_tensor(h_all, hblk.contiguous(), group=group)
    else:
        D_all, h_all = Dblk, hblk
    del Dblk
    mark("interface_all_gather")
    if D_all.numel() * 8 > (4 << 30) and D_all.is_cuda:
        torch.cuda.empty_cache()  # the root D needs one large block; give freed subtree buffers back first
    if root_mode == "factored":
        st.S_root_cols, g = ops.root_solve(D_all, h_all, Cpan, panels, plan.rank, plan.world, group, root_mode="factored")
    else:
        st.S_root_cols, g = ops.root_solve(D_all, h_all, Cpan, panels, plan.rank, plan.world, group)
    mark("root_solve")
    if marks:
        torch.cuda.synchronize()
        st.timing = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
    st.g_tilde_root = g
    st.multi = multi
    st.panels = panels
    st.col_index = ops.tensor(panel_column_index(panels, m)).to(torch.int64)
    return st


def root_S_action(st: ShardedState, plan: SubtreePlan, x, device=None, ops=None, group=None):
    """``S_root @ x`` for a boundary vector / matrix ``x`` in the root's face order, in either root mode (the
    quantity the reference stores as ``S_lst[-1]``, probed without forming it)."""
    ops = ops or CudaOps(device)
    xt = ops.tensor(x)
    xt = xt.reshape(xt.shape[0], -1)
    part = ops.matvec(st.S_root_cols, xt.index_select(0, st.col_index).contiguous())
    if plan.world > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    if st.root_mode == "factored":
        part = ops.root_apply(part.contiguous(), plan.rank, plan.world, group)
    return part


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
    part = ops.matvec(st.S_root_cols, g_ext.index_select(0, st.col_index).contiguous())
    if plan.world > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    if st.root_mode == "factored":  # part = -C g_ext: every rank applies D^-1 with its copy of the factors
        part = ops.root_apply(part.contiguous(), plan.rank, plan.world, group)
    g_int = part + gt
    kids = ops.root_scatter(g_ext.contiguous(), g_int.contiguous())
    mine = kids[plan.first_octant : plan.first_octant + plan.octants_per_rank]
    u = ops.down_local(mine, st.S_lst, st.g_tilde_lst, st.Y, st.v)
    return u[..., 0] if single else u
