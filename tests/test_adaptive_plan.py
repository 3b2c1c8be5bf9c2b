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
