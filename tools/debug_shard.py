import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import jaxhps_b200 as hps
from jaxhps_b200 import _dist_adaptive as da
from jaxhps_b200._adaptive_plan import get_plan
from oracle import hps_oracle_adaptive as ora
from test_oracle_adaptive import adaptive_problem
from adaptive_cases import boundary_fn
from _cases import rel_err

name = sys.argv[1] if len(sys.argv) > 1 else "adapt3d_p6q4"
case, dom, pb = adaptive_problem(name)
Yo, To, vo, ho = ora.local_solve_stage_adaptive_DtN(pb)
store = ora.merge_stage_adaptive_DtN(pb, To, ho)
ops = da.CudaAdaptiveOps("cuda")
shard = da.AdaptiveShardPlan(dom.root, 0, 1)
plan = get_plan(pb); rp = plan.by_id[id(dom.root)]
Ts, hs = [], []
for c in shard.children:
    sub = da.subtree_problem(pb, shard, c)
    T, h = ops.build_subtree(sub)
    kid = dom.root.children[c]
    rec = store[id(kid)]
    print("child", c, "leaf" if not kid.children else "node", "T err", rel_err(T.cpu().numpy(), rec["T"]), "h err", rel_err(h[:, 0].cpu().numpy(), rec["h"]), flush=True)
    T2, h2 = ops.compress(T, h, rp, c, pb.L_4f1 if not dom.bool_2D else pb.L_2f1, pb.L_1f4 if not dom.bool_2D else pb.L_1f2)
    Ts.append(T2); hs.append(h2)
S, g = ops.root_merge(Ts, hs, rp, 0, rp.ext_tbl.shape[0])
rr = store[id(dom.root)]
print("root S err", rel_err(S.cpu().numpy(), rr["S"]), "g err", rel_err(g[:, 0].cpu().numpy(), rr["g_tilde"]))
g_lst = dom.get_adaptive_boundary_data_lst(boundary_fn)
g_ext = np.concatenate(g_lst)
g_int = rr["S"] @ g_ext + rr["g_tilde"]
kids_o = ora.propagate_down_adaptive(dom.root, rr["S"], rr["g_tilde"], g_ext, pb.L_4f1 if not dom.bool_2D else pb.L_2f1, dom.q)
kids = ops.down_root(rp, ops.to_array(g_ext[:, None]), ops.to_array(g_int[:, None]), pb.L_4f1 if not dom.bool_2D else pb.L_2f1)
for c in range(len(kids)):
    print("kid g", c, rel_err(kids[c][:, 0].cpu().numpy(), kids_o[c]))
