"""world_size-2 (and 4) `gloo` runs of the subtree-sharded driver on CPU.  The collective /
partition logic is the product's (`jaxhps_b200/_dist.py`); the arithmetic is injected from the
CPU oracle so that the test needs no GPU.  The result must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxhps_b200 import _dist
from oracle import hps_oracle as orc
from _cases import make_domain, seeded_inputs


class OracleOps:
    """CPU test double for `_dist.CudaOps` built on the oracle (tests may import the oracle)."""

    def tensor(self, x):
        return torch.as_tensor(np.asarray(x), dtype=torch.float64)

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.float64)

    def local_solve(self, pb):
        return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in orc.local_solve_stage_uniform_3D_DtN(pb))

    def merge_subtrees(self, T, h, levels, n_roots):
        T, h = T.numpy(), h.numpy()
        S_lst, g_lst = [], []
        for _ in range(levels):
            n = T.shape[0] // 8
            outs = [orc.uniform_oct_merge_DtN(T[8 * i : 8 * i + 8], h[8 * i : 8 * i + 8]) for i in range(n)]
            S_lst.append(torch.from_numpy(np.stack([o[0] for o in outs])))
            T = np.stack([o[1] for o in outs])
            h = np.stack([o[2] for o in outs])
            g_lst.append(torch.from_numpy(np.stack([o[3] for o in outs])))
        assert T.shape[0] == n_roots
        return S_lst, g_lst, torch.from_numpy(T), torch.from_numpy(h)

    def root_pack(self, T_roots, h_roots, first_child):
        T, h = T_roots.numpy(), h_roots.numpy().reshape(T_roots.shape[0], T_roots.shape[1], -1)
        m = T.shape[-1] // 6
        D, C, hb = [], [], []
        for k in range(T.shape[0]):
            roles = orc._OCT_ROLES[first_child + k]
            int_faces = sorted((f for f in range(6) if roles[f][0] == "int"), key=lambda f: roles[f][1])
            ext_faces = [f for f in range(6) if roles[f][0] == "ext"]
            ii = np.concatenate([np.arange(f * m, (f + 1) * m) for f in int_faces])
            ee = np.concatenate([np.arange(f * m, (f + 1) * m) for f in ext_faces])
            D.append(T[k][np.ix_(ii, ii)]), C.append(T[k][np.ix_(ii, ee)]), hb.append(h[k][ii])
        return tuple(torch.from_numpy(np.ascontiguousarray(np.stack(x))) for x in (D, C, hb))

    def root_solve(self, Dblk_all, hblk_all, Cpan, panels, rank=0, world=1, group=None):
        Db, hb, Cp = Dblk_all.numpy(), hblk_all.numpy(), Cpan.numpy()
        m = Db.shape[-1] // 3
        D = np.zeros((12 * m, 12 * m))
        h_int = np.zeros((12 * m, hb.shape[-1]))
        C = np.zeros((12 * m, m * len(panels)))
        for c in range(8):
            roles = orc._OCT_ROLES[c]
            slots = sorted(roles[f][1] for f in range(6) if roles[f][0] == "int")
            for i, si in enumerate(slots):
                h_int[si * m : (si + 1) * m] += hb[c][i * m : (i + 1) * m]
                for j, sj in enumerate(slots):
                    D[si * m : (si + 1) * m, sj * m : (sj + 1) * m] += Db[c][i * m : (i + 1) * m, j * m : (j + 1) * m]
                for k, (pc, _) in enumerate(panels):  # panel k: m columns of child pc's C block
                    if pc == c:
                        C[si * m : (si + 1) * m, k * m : (k + 1) * m] = Cp[k][i * m : (i + 1) * m]
        D_inv = np.linalg.inv(D)
        return torch.from_numpy(-D_inv @ C), torch.from_numpy(-D_inv @ h_int)

    def matvec(self, S_cols, g_slice):
        return S_cols @ g_slice

    def root_scatter(self, g_ext, g_int):
        m = g_int.shape[0] // 12
        zero_S = np.zeros((12 * m, 24 * m))
        kids = [orc.propagate_down_oct_DtN(zero_S, g_ext[:, k].numpy(), g_int[:, k].numpy()) for k in range(g_ext.shape[1])]
        return torch.from_numpy(np.stack(kids, axis=-1))

    def down_local(self, g_roots, S_lst, g_lst, Y, v):
        bdry = g_roots.numpy()
        single = v.ndim == 2
        if single:
            bdry = bdry[..., 0]
        for level in range(len(S_lst) - 1, -1, -1):
            S, gt = S_lst[level].numpy(), g_lst[level].numpy()
            kids = [orc.propagate_down_oct_DtN(S[i], bdry[i], gt[i]) for i in range(bdry.shape[0])]
            bdry = np.concatenate(kids, axis=0)
        if single:
            u = np.einsum("ijk,ik->ij", Y.numpy(), bdry) + v.numpy()
            return torch.from_numpy(u[..., None])
        return torch.from_numpy(np.einsum("ijk,ikl->ijl", Y.numpy(), bdry) + v.numpy())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, p, q, L, nsrc, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        co, src, bdry = seeded_inputs(3, p, q, L, nsrc, seed)
        dom = make_domain(3, p, q, L)
        plan = _dist.SubtreePlan(L, rank, world)
        sl = plan.leaf_slice
        pb = _dist.local_problem(dom, plan, source=src[sl], **{k: v[sl] for k, v in co.items()})
        ops = OracleOps()
        st = _dist.build_solver_sharded(pb, plan, ops=ops)
        u = _dist.solve_sharded(pb, st, plan, bdry, ops=ops)
        np.save(os.path.join(out_dir, f"u_{rank}.npy"), u.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p,q,L,nsrc", [(2, 4, 2, 2, 1), (4, 4, 2, 1, 1), (2, 4, 2, 2, 2), (8, 4, 2, 2, 1)])
def test_sharded_build_and_solve_matches_single_process(tmp_path, world, p, q, L, nsrc):
    seed = 40 + world
    port = _free_port()
    mp.spawn(_worker, args=(world, port, p, q, L, nsrc, seed, str(tmp_path)), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / f"u_{r}.npy") for r in range(world)], axis=0)
    co, src, bdry = seeded_inputs(3, p, q, L, nsrc, seed)
    dom = make_domain(3, p, q, L)
    import jaxhps_b200 as hps

    pb = hps.PDEProblem(dom, source=src, **co)
    Y, T, v, h = orc.local_solve_stage_uniform_3D_DtN(pb)
    S, g = orc.merge_stage_uniform_3D_DtN(T, h, L)
    if nsrc > 1 and L == 1:
        pytest.skip("reference quirk: 3D multi-source needs L >= 2")
    ref = orc.down_pass_uniform_3D_DtN(bdry, S, g, Y, v)
    assert u.shape == ref.shape
    assert np.abs(u - ref).max() / np.abs(ref).max() < 1e-12


def test_subtree_plan_partitions_the_leaves():
    for world in (1, 2, 4, 8):
        slices = [_dist.SubtreePlan(3, r, world).leaf_slice for r in range(world)]
        covered = np.concatenate([np.arange(512)[s] for s in slices])
        assert np.array_equal(covered, np.arange(512))
    cols = np.concatenate([_dist.child_column_index(c, 1, 5) for c in range(8)])
    assert sorted(cols) == list(range(24 * 5))
    for world in (1, 2, 4, 8):  # the balanced panel assignment covers every exterior panel once, sorted inside a rank
        assign = _dist.balanced_panels(world)
        assert sorted(p for b in assign for p in b) == [(c, i) for c in range(8) for i in range(3)]
        assert all(len(b) == 24 // world for b in assign)
        for b in assign:
            first = [_dist._FIRST_SLOT[c] for c, _ in b]
            assert first == sorted(first)
        cols = np.concatenate([_dist.panel_column_index(b, 5) for b in assign])
        assert sorted(cols) == list(range(24 * 5))
        loads = [sum((1 - _dist._FIRST_SLOT[c] / 12) ** 2 for c, _ in b) for b in assign]
        assert max(loads) < 1.05 * 15.05 / world + 0.1  # forward-substitution weights: within 5 % of the mean
    assert np.array_equal(_dist.panel_column_index(_dist.balanced_panels(1)[0], 5), cols_ref := np.concatenate(
        [_dist.child_column_index(c, 1, 5) for c in range(8)]))
    with pytest.raises(ValueError):
        _dist.SubtreePlan(3, 0, 3)
