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


def seg_idx(start, width, rev, npp):
    idx = start + np.arange(width * npp)
    return idx[::-1] if rev else idx


def emu_compress(T, h, seg, npp, Lr, Lc):
    cols = [T[:, seg_idx(s, w, r, npp)] @ (Lr if w > 1 else np.eye(npp)) for s, w, r in seg]
    tmp = np.concatenate(cols, axis=1)
    rows = [(Lc if w > 1 else np.eye(npp)) @ tmp[seg_idx(s, w, r, npp)] for s, w, r in seg]
    hh = [(Lc if w > 1 else np.eye(npp)) @ h[seg_idx(s, w, r, npp)] for s, w, r in seg]
    return np.concatenate(rows, axis=0), np.concatenate(hh, axis=0)


def emu_merge(Tc, hc, plan):
    npp = plan.npp
    pan = lambda p: slice(p * npp, (p + 1) * npp)  # noqa: E731
    NI, NE = plan.int_tbl.shape[0], plan.ext_tbl.shape[0]
    owners = [((a, pa), (b, pb)) for a, pa, b, pb in plan.int_tbl] + [((c, p),) for c, p in plan.ext_tbl]
    n = (NI + NE) * npp
    M = np.zeros((n, n))
    rhs = np.zeros((n,) + hc[0].shape[1:])
    for I, own_r in enumerate(owners):
        for c, p in own_r:
            rhs[pan(I)] += hc[c][pan(p)]
        for J, own_c in enumerate(owners):
            for c, p in own_r:
                for c2, p2 in own_c:
                    if c == c2:
                        M[pan(I), pan(J)] += Tc[c][pan(p), pan(p2)]
    ni = NI * npp
    D, C, B, A = M[:ni, :ni], M[:ni, ni:], M[ni:, :ni], M[ni:, ni:]
    S = np.linalg.solve(D, -C)
    gt = np.linalg.solve(D, -rhs[:ni])
    # the non-zero blocks of B listed in bs_tbl must reproduce B S exactly as the dense product does
    BS = np.zeros_like(A)
    for c, r0, c0, Mb, Kb, s0, t0 in plan.bs_tbl:
        BS[t0 : t0 + Mb] += Tc[c][r0 : r0 + Mb, c0 : c0 + Kb] @ S[s0 : s0 + Kb]
    assert np.abs(BS - B @ S).max() <= 1e-12 * max(1.0, np.abs(BS).max())
    return S, A + B @ S, rhs[ni:] + B @ gt, gt


def emu_down(plan, S, gt, g_ext, Lr):
    npp = plan.npp
    NE = plan.ext_tbl.shape[0]
    g_all = np.concatenate([g_ext, S @ g_ext + gt])
    out = [np.full(ch.n, np.nan) for ch in plan.children]
    for c, sp, s, w, r in plan.down_tbl:
        panel = g_all[sp * npp : (sp + 1) * npp]
        out[c][seg_idx(s, w, r, npp)] = Lr @ panel if w > 1 else panel
    assert not any(np.isnan(o).any() for o in out)
    return out


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
