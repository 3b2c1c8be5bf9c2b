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
GPU_CASES = sorted(ADAPTIVE_CASES)


@pytest.mark.parametrize("name", GPU_CASES)
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


@pytest.mark.parametrize("name", ["adapt2d_p6q4_manual", "adapt3d_p4q2_manual", "adapt3d_p6q4"])
def test_blockwise_B_times_S_matches_dense(name, monkeypatch):
    """Large nodes form T = A + B S from the non-zero blocks of B only; force that path on small trees."""
    from jaxhps_b200 import adaptive

    G = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    monkeypatch.setattr(adaptive, "SPARSE_B_MIN_INTERFACE", 0)
    case, dom, pb = adaptive_problem(name)
    T_top = hps.build_solver(pb, return_top_T=True)
    rng = np.random.default_rng(case["seed"] + 1000)
    assert rel_err(T_top @ rng.normal(size=T_top.shape[1]), G["T_top_probe"]) < TOL
    for i, node in enumerate(internal_nodes(dom.root)):
        assert rel_err(node.data.h, G[f"h_{i}"]) < TOL and rel_err(node.data.g_tilde, G[f"g_tilde_{i}"]) < TOL
    assert rel_err(hps.solve(pb, dom.get_adaptive_boundary_data_lst(boundary_fn)), G["u"]) < TOL


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
    """BASELINE config 4 in small: the wavefront problem u = arctan(10 (|x + 0.05| - 0.7)) on an octree
    refined on its own source (reference `examples/wavefront_adaptive_discretization_3D.py:297-400`)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "run_config4", os.path.join(os.path.dirname(GOLDEN_DIR), "..", "tools", "run_config4.py"))
    cfg4 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg4)
    errs = []
    for tol in (1e-2, 1e-3):
        root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        dom = hps.Domain.from_adaptive_discretization(p=8, q=6, root=root, f=cfg4.source, tol=tol)
        assert not dom.bool_uniform and dom.n_leaves > 8
        one = np.ones(dom.interior_points.shape[:2])
        pb = hps.PDEProblem(dom, source=cfg4.source(dom.interior_points), D_xx_coefficients=one, D_yy_coefficients=one,
                            D_zz_coefficients=one)
        hps.build_solver(pb, host_device="cuda")
        u = hps.solve(pb, dom.get_adaptive_boundary_data_lst(cfg4.wavefront_soln))
        exact = cfg4.wavefront_soln(dom.interior_points)
        errs.append(np.abs(u - exact).max() / np.abs(exact).max())
    assert errs[0] < 5e-2 and errs[1] < errs[0], errs


@pytest.mark.parametrize("name", ["adapt3d_p6q4", "adapt2d_p8q6", "adapt3d_p4q2_manual"])
def test_sharded_driver_on_one_gpu_matches_reference(name):
    """`_dist_adaptive` with the CUDA ops and a single rank: root children built as independent subtrees,
    coarsened, merged through the column-window root merge, and solved — must reproduce the fixture."""
    import torch

    from jaxhps_b200 import _dist_adaptive as da

    G = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    case, dom, pb = adaptive_problem(name)
    ops = da.CudaAdaptiveOps(torch.device("cuda", torch.cuda.current_device()))
    st = da.build_solver_sharded_adaptive(pb, ops, rank=0, world=1)
    u, sl = da.solve_sharded_adaptive(st, dom.get_adaptive_boundary_data_lst(boundary_fn), ops)
    assert (sl.start, sl.stop) == (0, dom.n_leaves)
    assert rel_err(u[..., 0].cpu().numpy(), G["u"]) < TOL
    # the same through the distributed factorisation of the root system (single rank, forced)
    ops_d = da.CudaAdaptiveOps("cuda")
    ops_d.FORCE_DIST_LU, ops_d.DIST_LU_MIN_N = True, 0
    st0 = da.build_solver_sharded_adaptive(pb, ops_d, rank=0, world=1)
    u0, _ = da.solve_sharded_adaptive(st0, dom.get_adaptive_boundary_data_lst(boundary_fn), ops_d)
    assert rel_err(u0[..., 0].cpu().numpy(), G["u"]) < TOL
    full = st0["S_r"].cpu().numpy()
    i = internal_nodes(dom.root).index(dom.root)
    if f"S_{i}" in G:
        assert rel_err(full, G[f"S_{i}"]) < TOL


@pytest.mark.parametrize("dim,l2", [(2, False), (2, True), (3, False), (3, True)])
def test_mesh_generator_refinement_check_on_the_device(dim, l2):
    """`Domain.from_adaptive_discretization(..., device=cuda)` (criterion through ``hps_refine_check``) builds the same
    tree, leaf for leaf, as the host generator that `tests/test_adaptive_host.py` pins to the reference."""
    import jaxhps_b200 as hps
    from jaxhps_b200._tree import get_all_leaves

    def f(x):
        r2 = ((x - 0.3) ** 2).sum(-1)
        return np.exp(-40.0 * r2) + 0.2 * np.sin(3.0 * x[..., 0])

    def root():
        return hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0) if dim == 2 else hps.DiscretizationNode3D(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0)

    p, q = (8, 6) if dim == 2 else (6, 4)
    tol = 1e-4 if dim == 2 else 1e-2
    host = hps.Domain.from_adaptive_discretization(p, q, root(), f, tol, use_l_2_norm=l2)
    dev = hps.Domain.from_adaptive_discretization(p, q, root(), f, tol, use_l_2_norm=l2, device="cuda:0")
    lh, ld = get_all_leaves(host.root), get_all_leaves(dev.root)
    assert len(lh) == len(ld) and len(lh) > 2 ** dim
    assert np.array_equal(host.interior_points, dev.interior_points)
