"""Extended-precision arbitration of the ItI (complex128) path at BASELINE config 2's leaf conditioning.

The ItI leaf systems ``B = [G; A_interior]`` mix rows of very different size (impedance rows ~6e3, operator rows ~2e7 at
p=16, leaf side 1/32, k=100; cond(B) ~ 5e5), so two correct FP64 evaluations need not agree to 1e-10: LAPACK's pivoted
``inv`` (the reference's / the oracle's route) is itself 1e-12..1e-11 from the exact result at the leaves and
2e-11..4e-11 after two merge levels.  "Truth" here is the SAME algorithm — the oracle's code — evaluated in x87 extended
precision (eps 1.1e-19) on the same FP64 inputs (tests/_longdouble.py).  The bar: the CUDA path is at least as close to
the truth as the oracle is, for every operator of the build and for the solution, and within 1e-10 of the truth."""
import numpy as np
import pytest

from jaxhps_b200.down_pass import down_pass_uniform_2D_ItI
from jaxhps_b200.local_solve import local_solve_stage_uniform_2D_ItI
from jaxhps_b200.merge import merge_stage_uniform_2D_ItI
from oracle import hps_oracle as orc
from _cases import config2_problem, rel_err, seeded_problem
from _longdouble import iti_pipeline_truth

pytestmark = pytest.mark.gpu


def _all_outputs(ls, mg, dp, pb, L, bdry):
    Y, R, v, h = ls(pb)
    S, g, R_top = mg(R, h, L, return_T=True)
    u = dp(bdry, S, g, Y, v)
    names = ["Y", "R", "v", "h"] + [f"S_{i}" for i in range(L)] + [f"g_tilde_{i}" for i in range(L)] + ["R_top", "u"]
    return names, [np.asarray(a) for a in [Y, R, v, h] + list(S) + list(g) + [R_top, u]]


def _arbitrate(pb, L, bdry, label):
    t = iti_pipeline_truth(pb, L, bdry)
    truth = list(t[:4]) + t[4] + t[5] + [t[6], t[7]]
    names, gpu = _all_outputs(local_solve_stage_uniform_2D_ItI, merge_stage_uniform_2D_ItI, down_pass_uniform_2D_ItI, pb, L, bdry)
    _, ora = _all_outputs(orc.local_solve_stage_uniform_2D_ItI, orc.merge_stage_uniform_2D_ItI, orc.down_pass_uniform_2D_ItI, pb, L, bdry)
    eg = {n: rel_err(a.reshape(b.shape), b) for n, a, b in zip(names, gpu, truth)}
    eo = {n: rel_err(a.reshape(b.shape), b) for n, a, b in zip(names, ora, truth)}
    print(f"{label}: " + ", ".join(f"{n} gpu {eg[n]:.1e} / oracle {eo[n]:.1e}" for n in names))
    return eg, eo


@pytest.mark.parametrize("L", [2, 3])
def test_config2_leaf_physics_cuda_is_at_least_as_accurate_as_the_oracle(L):
    dom, pb, bdry = config2_problem(L, half_width=2**L / 64)
    eg, eo = _arbitrate(pb, L, bdry, f"ItI arbitration, config-2 leaves, L={L}")
    for n in eg:
        assert eg[n] < 1e-10, (n, eg[n])
        assert eg[n] <= max(eo[n], 1e-12), (n, eg[n], eo[n])


def test_seeded_p16_cuda_is_at_least_as_accurate_as_the_oracle():
    pb, bdry = seeded_problem(20, 16, 14, 2, 1, seed=116)
    eg, eo = _arbitrate(pb, 2, bdry, "ItI arbitration, seeded p=16 L=2")
    for n in eg:
        assert eg[n] < 1e-10, (n, eg[n])
        assert eg[n] <= max(eo[n], 1e-12), (n, eg[n], eo[n])
