"""world_size-2 / 4 `gloo` runs of the subtree-sharded ADAPTIVE driver on CPU.  Partitioning, exchange
and reduction logic is the product's (`jaxhps_b200/_dist_adaptive.py`); the arithmetic is injected from
the CPU oracle / the NumPy table interpreter, so no GPU is needed.  The result must equal the
single-process adaptive oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxhps_b200 import _dist_adaptive as da
from oracle import hps_oracle_adaptive as ora
from _adaptive_emu import emu_compress, emu_down, emu_merge
from test_oracle_adaptive import adaptive_problem

from adaptive_cases import boundary_fn


class OracleAdaptiveOps:
    """CPU test double of `_dist_adaptive.CudaAdaptiveOps`."""

    def to_array(self, x):
        return np.asarray(x, dtype=np.float64)

    def empty(self, shape):
        return np.empty(shape)

    def broadcast(self, arr, src):
        t = torch.from_numpy(arr)
        dist.broadcast(t, src=src)
        return arr

    def all_reduce(self, arr):
        arr = np.ascontiguousarray(arr)
        dist.all_reduce(torch.from_numpy(arr))
        return arr

    def build_subtree(self, sub):
        Y, T, v, h = ora.local_solve_stage_adaptive_DtN(sub)
        sub.Y, sub.v = Y, v
        if not sub.domain.root.children:
            return T[0], h[0][:, None]
        sub.store = ora.merge_stage_adaptive_DtN(sub, T, h)
        rec = sub.store[id(sub.domain.root)]
        return rec["T"], rec["h"][:, None]

    def compress(self, T, h, root_plan, c, L_refine, L_coarsen):
        ch = root_plan.children[c]
        if ch.identity:
            return np.ascontiguousarray(T), np.ascontiguousarray(h)
        T2, h2 = emu_compress(T, h, ch.seg, root_plan.npp, L_refine, L_coarsen)
        return np.ascontiguousarray(T2), np.ascontiguousarray(h2)

    def root_merge(self, Ts, hs, root_plan, e0, e1):
        S, _, _, gt = emu_merge(Ts, hs, root_plan)
        npp = root_plan.npp
        return np.ascontiguousarray(S[:, e0 * npp : e1 * npp]), gt

    def matvec(self, S, x):
        return S @ x

    def down_root(self, root_plan, g_ext, g_int, L_refine):
        zero_S = np.zeros((root_plan.n_int, root_plan.n_ext))
        cols = [emu_down(root_plan, zero_S, g_int[:, k], g_ext[:, k], L_refine) for k in range(g_ext.shape[1])]
        return [np.stack([col[c] for col in cols], axis=-1) for c in range(len(root_plan.children))]

    def down_subtree(self, sub, g):
        if not sub.domain.root.children:
            return (sub.Y[0] @ g[:, 0] + sub.v[0])[None, :, None]
        return ora.down_pass_adaptive_DtN(sub, sub.store, [g[:, 0]], sub.Y, sub.v)[..., None]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, dom, pb = adaptive_problem(name)
        ops = OracleAdaptiveOps()
        st = da.build_solver_sharded_adaptive(pb, ops)
        u, sl = da.solve_sharded_adaptive(st, dom.get_adaptive_boundary_data_lst(boundary_fn), ops)
        assert u.shape[0] == sl.stop - sl.start == st["shard"].leaves_per_rank[rank]
        np.save(os.path.join(out_dir, f"u_{rank}.npy"), u)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "adapt3d_p4q2_manual"), (4, "adapt3d_p4q2_l2"), (2, "adapt2d_p6q4_manual"),
                                        (4, "adapt2d_p8q6")])
def test_sharded_adaptive_matches_single_process_oracle(tmp_path, world, name):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / f"u_{r}.npy") for r in range(world)], axis=0)[..., 0]
    case, dom, pb = adaptive_problem(name)
    Y, T, v, h = ora.local_solve_stage_adaptive_DtN(pb)
    store = ora.merge_stage_adaptive_DtN(pb, T, h)
    ref = ora.down_pass_adaptive_DtN(pb, store, dom.get_adaptive_boundary_data_lst(boundary_fn), Y, v)
    assert u.shape == ref.shape
    assert np.abs(u - ref).max() / np.abs(ref).max() < 1e-11


def test_adaptive_shard_plan_partitions_leaves_and_columns():
    case, dom, pb = adaptive_problem("adapt3d_p6q4")
    for world in (1, 2, 4, 8):
        plans = [da.AdaptiveShardPlan(dom.root, r, world) for r in range(world)]
        covered = np.concatenate([np.arange(dom.n_leaves)[p.leaf_slice] for p in plans])
        assert np.array_equal(covered, np.arange(dom.n_leaves))
        assert sum(plans[0].leaves_per_rank) == dom.n_leaves
        cols = [plans[r].columns(51) for r in range(world)]
        assert cols[0][0] == 0 and cols[-1][1] == 51 and all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
    with pytest.raises(ValueError):
        da.AdaptiveShardPlan(dom.root, 0, 3)
