"""NumPy interpreter of the adaptive plan tables, written to mirror what the CUDA kernels of
csrc/adaptive.cu do with them (test infrastructure)."""
import numpy as np


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
