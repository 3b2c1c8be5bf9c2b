"""CPU oracle solution of BASELINE config 2 (2D variable-coefficient Helmholtz, ItI, p=16 q=14, L=6, k=100, gauss-bump
potential, complex128) -> tests/golden/config2_oracle_probe_L6.npz: strided probes of the solution, the
root g~ and the action of the root S and of the top-level R on fixed probe vectors."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _cases import config2_problem  # noqa: E402
from oracle import hps_oracle as orc  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
STRIDE = 101
dom, pb, g = config2_problem(L)
t0 = time.time()
Y, R, v, h = orc.local_solve_stage_uniform_2D_ItI(pb)
print("leaf stage", round(time.time() - t0, 1), "s", flush=True)
t0 = time.time()
S, gt, R_top = orc.merge_stage_uniform_2D_ItI(R, h, L, return_T=True)
print("merge stage", round(time.time() - t0, 1), "s", flush=True)
u = orc.down_pass_uniform_2D_ItI(g, S, gt, Y, v)
rng = np.random.default_rng(2000)
x = rng.normal(size=S[-1].shape[-1]) + 1j * rng.normal(size=S[-1].shape[-1])
np.savez_compressed(os.path.join(ROOT, f"tests/golden/config2_oracle_probe_L{L}.npz"), u_probe=u.reshape(-1)[::STRIDE], stride=STRIDE,
                    u_max=np.abs(u).max(), g_tilde_root=np.asarray(gt[-1]).reshape(-1), S_root_x=(np.asarray(S[-1])[0] @ x),
                    R_top_x=np.asarray(R_top) @ x, x=x, leaf_R_x=np.asarray(R[::64]) @ x[: R.shape[-1]], leaf_h=np.asarray(h[::64]))
print("done; max|u| =", np.abs(u).max(), flush=True)
