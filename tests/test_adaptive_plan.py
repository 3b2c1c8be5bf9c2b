"""The merge-plan tables of jaxhps_b200/_adaptive_plan.py, interpreted in NumPy exactly as the CUDA
kernels of csrc/adaptive.cu interpret them, must reproduce the adaptive oracle (CPU check of the host
logic; the kernels themselves are checked on the GPU in test_gpu_adaptive.py)."""
import numpy as np
import pytest

from jaxhps_b200._adaptive_plan import TreePlan
from oracle import hps_oracle_adaptive as ora
from _cases import rel_err
from test_oracle_adaptive import adaptive_problem

from adaptive_cases import ADAPTIVE_CASES, boundary_fn


from _adaptive_emu import emu_compress, emu_down, emu_merge  # noqa: E402


@pytest.mark.parametrize("name", sorted(ADAPTIVE_CASES))
def test_plan_tables_reproduce_the_oracle(name):
    case, dom, pb = adaptive_problem(name)
    Y, T, v, h = ora.local_solve_stage_adaptive_DtN(pb)
    store = ora.merge_stage_adaptive_DtN(pb, T, h)
    tp = TreePlan(dom.root, dom.q)
    Lr, Lc = (pb.L_2f1, pb.L_1f2) if dom.bool_2D else (pb.L_4f1, pb.L_1f4)
    assert [id(x) for x in tp.leaves] == [id(x) for x in ora._leaves(dom.root)]
    mine = {id(leaf): (T[i], h[i]) for i, leaf in enumerate(tp.leaves)}
    for plan in tp.nodes:  # deepest first
        Tc, hc = [], []
        for ch, kid in zip(plan.children, plan.node.children):
            Tk, hk = mine[id(kid)]
            assert Tk.shape[0] == ch.n == sum(tp.face_sizes(kid))
            if not ch.identity:
                Tk, hk = emu_compress(Tk, hk, ch.seg, plan.npp, Lr, Lc)
            assert Tk.shape[0] == ch.n_out
            Tc.append(Tk), hc.append(hk)
        S, Tn, hn, gt = emu_merge(Tc, hc, plan)
        rec = store[id(plan.node)]
        assert rel_err(S, rec["S"]) < 1e-10 and rel_err(Tn, rec["T"]) < 1e-10
        assert rel_err(hn, rec["h"]) < 1e-10 and rel_err(gt, rec["g_tilde"]) < 1e-10
        mine[id(plan.node)] = (Tn, hn)
        assert [getattr(plan.node, f"n_{f}") for f in range(tp.n_faces)] == tp.face_sizes(plan.node)
    # down pass through the tables
    g_lst = dom.get_adaptive_boundary_data_lst(boundary_fn)
    g_of = {id(dom.root): np.concatenate(g_lst)}
    for plan in reversed(tp.nodes):  # shallowest first
        rec = store[id(plan.node)]
        for kid, g in zip(plan.node.children, emu_down(plan, rec["S"], rec["g_tilde"], g_of[id(plan.node)], Lr)):
            g_of[id(kid)] = g
    u = np.stack([Y[i] @ g_of[id(leaf)] + v[i] for i, leaf in enumerate(tp.leaves)])
    assert rel_err(u, ora.down_pass_adaptive_DtN(pb, store, g_lst, Y, v)) < 1e-10
    packed = tp.pack()
    assert packed.dtype == np.int32 and all(("int" in p.off) != p.all_leaf_children for p in tp.nodes)


def _random_level_restricted_tree(dim, n_splits, seed, q):
    """Split random leaves, restoring the one-level balance after every split the way the mesh generator does."""
    import jaxhps_b200 as hps
    from jaxhps_b200._adaptive_discretization import _ensure_box, _position
    from jaxhps_b200._tree import _add_children, get_all_leaves

    rng = np.random.default_rng(seed)
    root = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0) if dim == 2 else hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    _add_children(root, root, q)
    for _ in range(n_splits):
        leaves = [leaf for leaf in get_all_leaves(root) if leaf.depth < 4]
        node = leaves[rng.integers(len(leaves))]
        _add_children(node, root, q)
        pending = [node]
        while pending:
            cur = pending.pop()
            pos, k = _position(root, cur), 1 << cur.depth
            for ax in range(dim):
                for step in (-1, 1):
                    nb = list(pos)
                    nb[ax] += step
                    if 0 <= nb[ax] < k:
                        made = _ensure_box(root, cur.depth, tuple(nb), q)
                        if made is not None:
                            pending.append(made)
    return root


@pytest.mark.parametrize("dim,seed,n_splits", [(2, 1, 6), (2, 2, 10), (2, 3, 14), (3, 4, 2), (3, 5, 3), (3, 6, 4)])
def test_plan_tables_on_random_level_restricted_trees(dim, seed, n_splits):
    """Random trees exercise table shapes the fixtures do not: several coarsened runs on one face, level jumps on
    both sides of one interface, jumps at interfaces high up in the tree."""
    import jaxhps_b200 as hps

    p, q = (6, 4) if dim == 2 else (4, 2)
    root = _random_level_restricted_tree(dim, n_splits, seed, q)
    dom = hps.Domain(p=p, q=q, root=root)
    rng = np.random.default_rng(100 + seed)
    shp = dom.interior_points.shape[:2]
    co = {"D_xx_coefficients": 1 + 0.1 * rng.normal(size=shp), "D_yy_coefficients": np.ones(shp), "I_coefficients": rng.normal(size=shp)}
    if dim == 3:
        co["D_zz_coefficients"] = 1 + 0.1 * rng.normal(size=shp)
    pb = hps.PDEProblem(dom, source=rng.normal(size=shp), **co)
    Y, T, v, h = ora.local_solve_stage_adaptive_DtN(pb)
    store = ora.merge_stage_adaptive_DtN(pb, T, h)
    tp = TreePlan(root, q)
    Lr, Lc = (pb.L_2f1, pb.L_1f2) if dim == 2 else (pb.L_4f1, pb.L_1f4)
    mine = {id(leaf): (T[i], h[i]) for i, leaf in enumerate(tp.leaves)}
    n_coarsened = 0
    for plan in tp.nodes:
        Tc, hc = [], []
        for ch, kid in zip(plan.children, plan.node.children):
            Tk, hk = mine[id(kid)]
            n_coarsened += int((ch.seg[:, 1] > 1).sum())
            if not ch.identity:
                Tk, hk = emu_compress(Tk, hk, ch.seg, plan.npp, Lr, Lc)
            Tc.append(Tk), hc.append(hk)
        S, Tn, hn, gt = emu_merge(Tc, hc, plan)
        rec = store[id(plan.node)]
        assert rel_err(S, rec["S"]) < 1e-9 and rel_err(Tn, rec["T"]) < 1e-9 and rel_err(gt, rec["g_tilde"]) < 1e-9
        mine[id(plan.node)] = (Tn, hn)
        assert [getattr(plan.node, f"n_{f}") for f in range(tp.n_faces)] == tp.face_sizes(plan.node)
    assert n_coarsened > 0, "the random tree should contain level jumps"
    g_root = rng.normal(size=tp.n_points(root))
    g_of = {id(root): g_root}
    for plan in reversed(tp.nodes):
        rec = store[id(plan.node)]
        for kid, g in zip(plan.node.children, emu_down(plan, rec["S"], rec["g_tilde"], g_of[id(plan.node)], Lr)):
            g_of[id(kid)] = g
    u = np.stack([Y[i] @ g_of[id(leaf)] + v[i] for i, leaf in enumerate(tp.leaves)])
    sizes = tp.face_sizes(root)
    g_lst = np.split(g_root, np.cumsum(sizes)[:-1])
    assert rel_err(u, ora.down_pass_adaptive_DtN(pb, store, g_lst, Y, v)) < 1e-9
