"""BASELINE config 3 (3D, p=12, q=10, DtN, FP64) at L=2 and L=3: every stage of the CUDA path against the CPU oracle
through the committed probe fixtures (tools/oracle_config3.py -> tests/golden/config3_oracle_probe_L{2,3}.npz).
Operators are compared through their action on fixed probe vectors: leaf Y, T, v, h; every level's S, g~, T, h
(incl. the m=400 cluster-sized merges and the m=1600 root merge, n_int = 19 200); T_top; the solution u.
Tolerance 1e-10 relative (max-norm), the north_star bar."""
import os

import numpy as np
import pytest
import torch

from jaxhps_b200.down_pass import down_pass_uniform_3D_DtN
from jaxhps_b200.local_solve import local_solve_stage_uniform_3D_DtN
from jaxhps_b200.merge import _merge_stage
from _cases import GOLDEN_DIR, config3_probe, config3_problem

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rel(a, b):
    b = torch.as_tensor(np.asarray(b), device=a.device)
    assert tuple(a.shape) == tuple(b.shape), (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("L", [2, 3])
def test_config3_every_stage_matches_the_oracle_fixture(L):
    path = os.path.join(GOLDEN_DIR, f"config3_oracle_probe_L{L}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tools/oracle_config3.py {L})")
    G = dict(np.load(path))
    _, leaf_stride, u_stride = (int(t) for t in G["meta"])
    dev = torch.device("cuda:0")
    pb, bdry = config3_problem(L)
    Y, T, v, h = local_solve_stage_uniform_3D_DtN(pb, device=dev, host_device=dev)
    x = torch.as_tensor(config3_probe(T.shape[-1], 0), device=dev)
    errs = {"leaf_T_x": _rel(T @ x, G["leaf_T_x"]), "leaf_h": _rel(h, G["leaf_h"]),
            "leaf_Y_x": _rel(Y[::leaf_stride] @ x, G["leaf_Y_x"]), "leaf_v": _rel(v[::leaf_stride], G["leaf_v"])}
    # level by level, so that every intermediate T is seen (the stage function only returns the last one)
    S_lst, g_lst = [], []
    T_cur, h_cur = T, h
    for level in range(L, 0, -1):
        k = L - level
        S1, g1, T_cur, h_cur = _merge_stage(T_cur, h_cur, 1, 3, dev, dev, True, False, n_roots=T_cur.shape[0] // 8)
        S, g = S1[0], g1[0]
        xk = torch.as_tensor(config3_probe(S.shape[-1], level), device=dev)
        errs[f"S_x_{k}"] = _rel(S @ xk, G[f"S_x_{k}"])
        errs[f"g_tilde_{k}"] = _rel(g, G[f"g_tilde_{k}"])
        errs[f"T_x_{k}"] = _rel(T_cur @ xk, G[f"T_x_{k}"])
        errs[f"h_{k}"] = _rel(h_cur, G[f"h_{k}"])
        S_lst.append(S if level > 1 else S[0])
        g_lst.append(g if level > 1 else g[0])
    del T_cur
    u = down_pass_uniform_3D_DtN(torch.as_tensor(bdry, device=dev), S_lst, g_lst, Y, v, device=dev, host_device=dev)
    errs["u"] = float((u.reshape(-1)[::u_stride] - torch.as_tensor(G["u_probe"], device=dev)).abs().max() / float(G["u_max"]))
    print(f"config 3 L={L}: " + ", ".join(f"{k} {e:.1e}" for k, e in errs.items()))
    bad = {k: e for k, e in errs.items() if not e < TOL}
    assert not bad, bad


@pytest.mark.parametrize("root_mode", ["S", "factored"])
@pytest.mark.parametrize("L", [2, 3])
def test_config3_sharded_driver_root_modes(L, root_mode):
    """The sharded driver on one rank (root merge through hps_lu_dist_run on the P2P segment) in both root modes:
    ``S`` (root S formed) and ``factored`` (S-free: LU of D kept, D^-1 applied per solve).  The action of the root S
    on the fixture's probe vector, the root g~ and the solution must match the oracle fixture at 1e-10."""
    from jaxhps_b200 import _dist

    path = os.path.join(GOLDEN_DIR, f"config3_oracle_probe_L{L}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    G = dict(np.load(path))
    u_stride = int(G["meta"][2])
    dev = torch.device("cuda:0")
    pb_full, bdry = config3_problem(L)
    plan = _dist.SubtreePlan(L, 0, 1)
    co = {k: getattr(pb_full, k) for k in ("D_xx_coefficients", "D_yy_coefficients", "D_zz_coefficients", "D_x_coefficients",
                                           "I_coefficients")}
    pb = _dist.local_problem(pb_full.domain, plan, source=pb_full.source, **co)
    ops = _dist.CudaOps(dev)
    ops.FORCE_DIST_LU, ops.DIST_LU_MIN_N = True, 0
    st = _dist.build_solver_sharded(pb, plan, ops=ops, root_mode=root_mode)
    k = L - 1
    x = config3_probe(G[f"S_x_{k}"].shape[-1] * 2, 1)  # the root level's probe (salt = level 1), length 24 m
    Sx = _dist.root_S_action(st, plan, x, ops=ops)[:, 0]
    errs = {"S_x_root": _rel(Sx, G[f"S_x_{k}"][0]), "g_tilde_root": _rel(st.g_tilde_root[:, 0], G[f"g_tilde_{k}"][0])}
    u = _dist.solve_sharded(pb, st, plan, bdry, ops=ops)
    errs["u"] = float((u.reshape(-1)[::u_stride] - torch.as_tensor(G["u_probe"], device=dev)).abs().max() / float(G["u_max"]))
    print(f"config 3 L={L} sharded root_mode={root_mode}: " + ", ".join(f"{a} {e:.1e}" for a, e in errs.items()))
    assert all(e < TOL for e in errs.values()), errs
