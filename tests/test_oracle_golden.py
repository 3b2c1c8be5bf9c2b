"""Pin the NumPy oracle (and the host pre-compute) to outputs of the REFERENCE ITSELF: the
fixtures in tests/golden/*.npz were produced by running /root/reference's own code on the
NumPy `jax` shim (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import hps_oracle as orc
from _cases import golden_names, load_golden, rel_err, seeded_problem

TOL = 1e-11


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(name):
    G = load_golden(name)
    dim, p, q, L, nsrc, seed = (int(x) for x in G["meta"])
    pb, bdry = seeded_problem(dim, p, q, L, nsrc, seed)
    if dim == 20:
        ls, mg, dp = orc.local_solve_stage_uniform_2D_ItI, orc.merge_stage_uniform_2D_ItI, orc.down_pass_uniform_2D_ItI
    elif dim == 3:
        ls, mg, dp = orc.local_solve_stage_uniform_3D_DtN, orc.merge_stage_uniform_3D_DtN, orc.down_pass_uniform_3D_DtN
    else:
        ls, mg, dp = orc.local_solve_stage_uniform_2D_DtN, orc.merge_stage_uniform_2D_DtN, orc.down_pass_uniform_2D_DtN
    Y, T, v, h = ls(pb)
    S_lst, g_lst, T_top = mg(T, h, L, return_T=True)
    u = dp(bdry, S_lst, g_lst, Y, v)
    assert rel_err(v, G["v"]) < TOL and rel_err(h, G["h"]) < TOL
    for i, g in enumerate(g_lst):
        assert rel_err(g, G[f"g_tilde_{i}"]) < TOL
    probe = np.random.default_rng(seed + 1000).normal(size=T_top.shape[1])
    assert rel_err(T_top @ probe, G["T_top_probe"]) < TOL
    assert rel_err(u, G["u"]) < TOL
    if "Y" in G:
        assert rel_err(Y, G["Y"]) < TOL and rel_err(T, G["T"]) < TOL and rel_err(T_top, G["T_top"]) < TOL
        for i, S in enumerate(S_lst):
            assert rel_err(S, G[f"S_{i}"]) < TOL
        # host pre-compute against the reference's operators and point clouds
        assert rel_err(pb.P, G["P"]) < 1e-14 and rel_err(pb.D_x, G["D_x"]) < 1e-14
        if dim == 20:
            assert rel_err(pb.G, G["G"]) < 1e-14 and rel_err(pb.QH, G["QH"]) < 1e-14
        else:
            assert rel_err(pb.Q, G["Q"]) < 1e-14
        assert rel_err(pb.domain.interior_points, G["interior_points"]) < 1e-15
        assert rel_err(pb.domain.boundary_points, G["boundary_points"]) < 1e-15


def test_root_entries_follow_reference_shapes():
    """3D stores the root S / g_tilde without a batch axis, 2D with a leading 1."""
    pb, _ = seeded_problem(3, 4, 2, 2, seed=1)
    Y, T, v, h = orc.local_solve_stage_uniform_3D_DtN(pb)
    S, g = orc.merge_stage_uniform_3D_DtN(T, h, 2)
    assert S[0].shape == (8, 48, 96) and S[1].shape == (192, 384) and g[1].shape == (192,)
    pb, _ = seeded_problem(2, 6, 4, 2, seed=1)
    Y, T, v, h = orc.local_solve_stage_uniform_2D_DtN(pb)
    S, g = orc.merge_stage_uniform_2D_DtN(T, h, 2)
    assert S[0].shape == (4, 16, 32) and S[1].shape == (1, 32, 64) and g[1].shape == (1, 32)
