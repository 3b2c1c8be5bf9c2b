"""ItI leaf accuracy study (developer tool): GPU vs oracle vs extended-precision truth on seeded problems."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from _cases import rel_err, seeded_problem  # noqa: E402
from _longdouble import iti_leaf_truth  # noqa: E402
from jaxhps_b200.local_solve import local_solve_stage_uniform_2D_ItI  # noqa: E402
from jaxhps_b200.merge import merge_stage_uniform_2D_ItI  # noqa: E402
from oracle import hps_oracle as orc  # noqa: E402

for (p, q, L, nsrc) in [(6, 4, 2, 1), (8, 6, 3, 2), (16, 14, 3, 1)]:
    pb, bdry = seeded_problem(20, p, q, L, nsrc, seed=100 + p)
    Yo, Ro, vo, ho = orc.local_solve_stage_uniform_2D_ItI(pb)
    Y, R, v, h = local_solve_stage_uniform_2D_ItI(pb)
    print(f"p={p} q={q} L={L}: GPU vs oracle  Y {rel_err(Y, Yo):.2e} R {rel_err(R, Ro):.2e} v {rel_err(v, vo):.2e} h {rel_err(h, ho):.2e}")
    for leaf in (0, 4**L // 2, 4**L - 1):
        Yt, Rt, vt, ht = iti_leaf_truth(pb, leaf)
        sq = (lambda a: a[leaf][..., None] if a[leaf].ndim == 1 else a[leaf])
        eg = [rel_err(Y[leaf], Yt), rel_err(R[leaf], Rt), rel_err(sq(v), vt), rel_err(sq(h), ht)]
        eo = [rel_err(Yo[leaf], Yt), rel_err(Ro[leaf], Rt), rel_err(sq(vo), vt), rel_err(sq(ho), ht)]
        print(f"   leaf {leaf}: GPU vs truth " + " ".join(f"{e:.2e}" for e in eg) + "   oracle vs truth " + " ".join(f"{e:.2e}" for e in eo))
    So, go, Tto = orc.merge_stage_uniform_2D_ItI(Ro, ho, L, return_T=True)
    S, g, Tt = merge_stage_uniform_2D_ItI(Ro, ho, L, return_T=True)
    print("   merges (same oracle leaf inputs): " + " ".join(f"{rel_err(a, b):.2e}" for a, b in zip(S + g + [Tt], So + go + [Tto])))
