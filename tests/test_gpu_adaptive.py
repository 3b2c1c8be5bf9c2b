"""Adaptive (non-uniform tree) path on the GPU against the adaptive oracle and the reference fixtures."""
import os

import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200.adaptive import (
    down_pass_adaptive_2D_DtN,
    down_pass_adaptive_3D_DtN,
    local_solve_stage_adaptive_2D_DtN,
    local_solve_stage_adaptive_3D_DtN,
    merge_stage_adaptive_2D_DtN,
    merge_stage_adaptive_3D_DtN,
)
from oracle import hps_oracle_adaptive as ora
from _cases import GOLDEN_DIR, rel_err
from test_oracle_adaptive import adaptive_problem

from adaptive_cases import ADAPTIVE_CASES, boundary_fn, internal_nodes

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("name", sorted(ADAPTIVE_CASES))
def test_adaptive_stages_match_oracle_and_reference(name):
    G = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    case, dom, pb = adaptive_problem(name)
    two_d = dom.bool_2D
    ls = local_solve_stage_adaptive_2D_DtN if two_d else local_solve_stage_adaptive_3D_DtN
    mg = merge_stage_adaptive_2D_DtN if two_d else merge_stage_adaptive_3D_DtN
    dp = down_pass_adaptive_2D_DtN if two_d else down_pass_adaptive_3D_DtN
    Yo, To, vo, ho = ora.local_solve_stage_adaptive_DtN(pb)
    Y, T, v, h = ls(pb)
    for a, b in ((Y, Yo), (T, To), (v, vo), (h, ho)):
        assert rel_err(a, b) < TOL
    store = ora.merge_stage_adaptive_DtN(pb, To, ho)
    n_root = mg(pb, T, h)
    assert n_root == store[id(dom.root)]["S"].shape[0]
    for i, node in enumerate(internal_nodes(dom.root)):
        rec = store[id(node)]
        assert rel_err(node.data.S, rec["S"]) < TOL, (i, node)
        assert rel_err(node.data.g_tilde, rec["g_tilde"]) < TOL and rel_err(node.data.h, rec["h"]) < TOL
        assert rel_err(node.data.g_tilde, G[f"g_tilde_{i}"]) < TOL
    assert rel_err(dom.root.data.T, store[id(dom.root)]["T"]) < TOL
    rng = np.random.default_rng(case["seed"] + 1000)
    assert rel_err(dom.root.data.T @ rng.normal(size=dom.root.data.T.shape[1]), G["T_top_probe"]) < TOL
    g_lst = dom.get_adaptive_boundary_data_lst(boundary_fn)
    pb.Y, pb.v = Y, v
    u = dp(pb, g_lst)
    assert rel_err(u, ora.down_pass_adaptive_DtN(pb, store, g_lst, Yo, vo)) < TOL
    assert rel_err(u, G["u"]) < TOL


@pytest.mark.parametrize("name", ["adapt2d_p8q6", "adapt3d_p6q4"])
def test_adaptive_build_solver_and_solve(name):
    """Public API: build_solver / solve on an adaptive Domain, operators resident on the device."""
    G = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    case, dom, pb = adaptive_problem(name)
    T_top = hps.build_solver(pb, return_top_T=True, host_device="cuda")
    assert T_top.shape[0] == dom.boundary_points.shape[0]
    g_lst = dom.get_adaptive_boundary_data_lst(boundary_fn)
    u = hps.solve(pb, g_lst)
    assert rel_err(u, G["u"]) < TOL
    with pytest.raises(ValueError):
        hps.solve(pb, np.concatenate(g_lst))
    # second solve with other data reuses the factorisation
    u2 = hps.solve(pb, [2 * g for g in g_lst])
    case2, dom2, pb2 = adaptive_problem(name)
    pb2.source = np.zeros_like(pb2.source)
    hps.build_solver(pb2)
    u_h = hps.solve(pb2, g_lst)  # homogeneous part
    assert rel_err(u2 - u, u_h) < 1e-9


def test_wavefront_adaptive_3d_reaches_analytic_solution():
    """Config 4 in small: adaptive octree refined on the source of a manufactured solution."""
    def u_true(x):
        return np.exp(-20 * ((x[..., 0] - 0.4) ** 2 + (x[..., 1] - 0.5) ** 2 + (x[..., 2] - 0.6) ** 2))

    def lap(x):
        r2 = (x[..., 0] - 0.4) ** 2 + (x[..., 1] - 0.5) ** 2 + (x[..., 2] - 0.6) ** 2
        return (1600 * r2 - 120) * np.exp(-20 * r2)

    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    dom = hps.Domain.from_adaptive_discretization(p=8, q=6, root=root, f=lap, tol=1e-2)
    assert not dom.bool_uniform and dom.n_leaves > 8
    one = np.ones(dom.interior_points.shape[:2])
    pb = hps.PDEProblem(dom, source=lap(dom.interior_points), D_xx_coefficients=one, D_yy_coefficients=one, D_zz_coefficients=one)
    hps.build_solver(pb, host_device="cuda")
    u = hps.solve(pb, dom.get_adaptive_boundary_data_lst(u_true))
    err = np.abs(u - u_true(dom.interior_points)).max()
    assert err < 5e-3, err
