"""CPU oracle run of BASELINE config 3 (3D, p=12, q=10, DtN, FP64) at L=2 or L=3 on the seeded parity problem
of tests/_cases.config3_problem; writes probe fixtures tests/golden/config3_oracle_probe_L{L}.npz:

  leaf level : T @ x, h for every leaf; Y @ x and v for every 16th leaf
  level k    : S_k @ x, g~_k, T_k @ x, h_k for every merge (the root: S @ x, g~, T_top @ x)
  solution   : every 13th value of u

Operators are stored through their action on fixed probe vectors (tests/_cases.config3_probe) so that the
fixtures stay small; the GPU tests (tests/test_gpu_config3.py) compare at 1e-10.
usage: python tools/oracle_config3.py L      (L=3: ~20-40 min on 8 cores, ~40 GB of RAM)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _cases import config3_probe, config3_problem  # noqa: E402
from oracle import hps_oracle as orc  # noqa: E402

L = int(sys.argv[1])
LEAF_STRIDE, U_STRIDE = 16, 13
pb, bdry = config3_problem(L)
out = {"meta": np.array([L, LEAF_STRIDE, U_STRIDE])}
t0 = time.time()
Y, T, v, h = orc.local_solve_stage_uniform_3D_DtN(pb)
print(f"leaf stage {time.time() - t0:.1f} s", flush=True)
x = config3_probe(T.shape[-1], 0)
out["leaf_T_x"] = T @ x
out["leaf_h"] = h
out["leaf_Y_x"] = Y[::LEAF_STRIDE] @ x
out["leaf_v"] = v[::LEAF_STRIDE]

S_lst, g_lst = [], []
T_arr, h_arr = T, h
for level in range(L, 0, -1):
    t0 = time.time()
    n = T_arr.shape[0] // 8
    m = T_arr.shape[-1] // 6
    x = config3_probe(24 * m, level)
    last = level == 1
    Ss, Ts, hs, gs, Sx, Tx = [], [], [], [], [], []
    for i in range(n):
        if last:
            S, Tp, ho, g = orc.uniform_oct_merge_DtN(T_arr[8 * i : 8 * i + 8], h_arr[8 * i : 8 * i + 8], need_T=False, probe=x)
            Tx.append(Tp)
        else:
            S, Tm, ho, g = orc.uniform_oct_merge_DtN(T_arr[8 * i : 8 * i + 8], h_arr[8 * i : 8 * i + 8])
            Ts.append(Tm)
            Tx.append(Tm @ x)
        Ss.append(S), hs.append(ho), gs.append(g), Sx.append(S @ x)
    k = L - level
    out[f"S_x_{k}"], out[f"g_tilde_{k}"] = np.stack(Sx), np.stack(gs)
    out[f"T_x_{k}"], out[f"h_{k}"] = np.stack(Tx), np.stack(hs)
    S_lst.append(np.stack(Ss) if not last else Ss[0])
    g_lst.append(np.stack(gs) if not last else gs[0])
    if not last:
        T_arr, h_arr = np.stack(Ts), np.stack(hs)
    del Ss, Ts
    print(f"merge level {level} (m={m}, {n} merges) {time.time() - t0:.1f} s", flush=True)
del T_arr
u = orc.down_pass_uniform_3D_DtN(bdry, S_lst, g_lst, Y, v)
out["u_probe"] = u.reshape(-1)[::U_STRIDE]
out["u_max"] = np.abs(u).max()
path = os.path.join(ROOT, "tests", "golden", f"config3_oracle_probe_L{L}.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) / 1e6, "MB", flush=True)
