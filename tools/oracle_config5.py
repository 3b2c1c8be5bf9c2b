"""CPU oracle solution of BASELINE config 5 (see tools/run_config5.py) on the stored octree; writes a
strided probe of the solution to tests/golden/config5_oracle_probe_p10.npz so that the GPU run at the full
configuration can be checked against the (reference-pinned) oracle.  Takes several minutes of CPU time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import jaxhps_b200 as hps  # noqa: E402
import run_config5 as c5  # noqa: E402
from oracle import hps_oracle_adaptive as ora  # noqa: E402

P = 10
STRIDE = 997
root = hps.DiscretizationNode3D(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0)
dom = hps.Domain(p=P, q=P - 2, root=c5.decode_tree(root, np.load(os.path.join(ROOT, "tools/data/config5_tree_p10_tol1e-3.npy")), P - 2))
pb = c5.build_problem(dom)
t0 = time.time()
Y, T, v, h = ora.local_solve_stage_adaptive_DtN(pb)
print("leaf stage", round(time.time() - t0, 1), "s", flush=True)
t0 = time.time()
store = ora.merge_stage_adaptive_DtN(pb, T, h)
print("merge stage", round(time.time() - t0, 1), "s", flush=True)
g = dom.get_adaptive_boundary_data_lst(lambda x: np.zeros(x.shape[:-1]))
u = ora.down_pass_adaptive_DtN(pb, store, g, Y, v)
np.savez_compressed(os.path.join(ROOT, "tests/golden/config5_oracle_probe_p10.npz"), u_probe=u.reshape(-1)[::STRIDE],
                    stride=STRIDE, n_leaves=dom.n_leaves, u_max=np.abs(u).max())
print("done; max|u| =", np.abs(u).max(), flush=True)
