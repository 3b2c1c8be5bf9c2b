"""Pin the adaptive NumPy oracle to outputs of the reference itself (tests/golden/adapt_*.npz,
produced by tests/golden/make_golden_adaptive.py on the NumPy `jax` shim)."""
import os
import sys

import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200._tree import add_eight_children, add_four_children
from oracle import hps_oracle_adaptive as ora
from _cases import GOLDEN_DIR, rel_err

sys.path.insert(0, GOLDEN_DIR)
from adaptive_cases import ADAPTIVE_CASES, boundary_fn, build_domain, internal_nodes, seeded_fields  # noqa: E402

TOL = 1e-11


def adaptive_problem(name):
    case = ADAPTIVE_CASES[name]
    add = add_four_children if case["dim"] == 2 else add_eight_children
    dom = build_domain(hps, lambda n, r, q: add(n, root=r, q=q), case)
    co, src = seeded_fields(case, dom.n_leaves)
    return case, dom, hps.PDEProblem(dom, source=src, **co)


@pytest.mark.parametrize("name", sorted(ADAPTIVE_CASES))
def test_adaptive_oracle_matches_reference_fixture(name):
    G = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    case, dom, pb = adaptive_problem(name)
    Y, T, v, h = ora.local_solve_stage_adaptive_DtN(pb)
    assert rel_err(v, G["v"]) < TOL
    store = ora.merge_stage_adaptive_DtN(pb, T, h)
    g_lst = dom.get_adaptive_boundary_data_lst(boundary_fn)
    assert rel_err(np.concatenate(g_lst), G["g_bdry"]) < 1e-15
    u = ora.down_pass_adaptive_DtN(pb, store, g_lst, Y, v)
    rng = np.random.default_rng(case["seed"] + 1000)
    T_top = store[id(dom.root)]["T"]
    assert rel_err(T_top @ rng.normal(size=T_top.shape[1]), G["T_top_probe"]) < TOL
    for i, node in enumerate(internal_nodes(dom.root)):
        rec = store[id(node)]
        assert rel_err(rec["g_tilde"], G[f"g_tilde_{i}"]) < TOL and rel_err(rec["h"], G[f"h_{i}"]) < TOL
        assert rel_err(rec["S"] @ rng.normal(size=rec["S"].shape[1]), G[f"S_probe_{i}"]) < TOL
        if f"S_{i}" in G:
            assert rel_err(rec["S"], G[f"S_{i}"]) < TOL
    assert rel_err(u, G["u"]) < TOL
    if "T_top" in G:
        assert rel_err(T_top, G["T_top"]) < TOL and rel_err(Y, G["Y"]) < TOL
        assert rel_err(T, G["T_leaf"]) < TOL and rel_err(h, G["h_leaf"]) < TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_adaptive_oracle_reproduces_polynomial_solution(dim):
    """Known answer on a non-uniform tree: a quadratic solution is represented exactly by every leaf and by
    the Gauss panels, and the 4-to-1 / 2-to-1 projections are exact for it, so the adaptive build + solve
    must return it to rounding (the uniform analogue is the reference's `test_single_merge_accuracy.py`)."""
    if dim == 2:
        root = hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0)
        add, p, q = add_four_children, 8, 6
        paths = [[], [1], [1, 1], [3]]
        u = lambda x: x[..., 0] ** 2 - x[..., 1] ** 2 + 0.5 * x[..., 0] * x[..., 1]  # noqa: E731
        lap = lambda x: 3.0 * 2 * np.ones(x.shape[:-1])  # noqa: E731  u2 = u + 1.5 (x^2 + y^2)
        u2 = lambda x: u(x) + 1.5 * (x[..., 0] ** 2 + x[..., 1] ** 2)  # noqa: E731
    else:
        root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        add, p, q = add_eight_children, 6, 4
        paths = [[], [6], [6, 6], [0]]
        u = lambda x: x[..., 0] ** 2 - x[..., 2] ** 2 + x[..., 1] * x[..., 2]  # noqa: E731
        lap = lambda x: 6.0 * np.ones(x.shape[:-1])  # noqa: E731
        u2 = lambda x: u(x) + x[..., 0] ** 2 + x[..., 1] ** 2 + x[..., 2] ** 2  # noqa: E731
    for path in paths:
        node = root
        for c in path:
            node = node.children[c]
        add(node, root=root, q=q)
    dom = hps.Domain(p=p, q=q, root=root)
    one = np.ones(dom.interior_points.shape[:2])
    co = {"D_xx_coefficients": one, "D_yy_coefficients": one}
    if dim == 3:
        co["D_zz_coefficients"] = one
    pb = hps.PDEProblem(dom, source=lap(dom.interior_points), **co)
    Y, T, v, h = ora.local_solve_stage_adaptive_DtN(pb)
    store = ora.merge_stage_adaptive_DtN(pb, T, h)
    sol = ora.down_pass_adaptive_DtN(pb, store, dom.get_adaptive_boundary_data_lst(u2), Y, v)
    assert np.abs(sol - u2(dom.interior_points)).max() < 1e-10
    # the root DtN map applied to the trace of the harmonic part returns its normal derivative on every
    # boundary panel: check through T g = du/dn summed against the Gauss weights (net flux of a harmonic function = 0)
    g_h = np.concatenate(dom.get_adaptive_boundary_data_lst(u))
    flux = store[id(root)]["T"] @ g_h
    sizes = ora._face_index_ranges(root, dim)
    # every leaf face panel has the same Gauss weights up to its area
    w1 = np.polynomial.legendre.leggauss(q)[1]
    total = 0.0
    for f, leaves in enumerate([ora._face_leaves(root, f) for f in range(2 * dim)]):
        at = sizes[f][0]
        for leaf in leaves:
            side = leaf.xmax - leaf.xmin
            w = (w1 * side / 2) if dim == 2 else np.outer(w1, w1).reshape(-1) * (side / 2) ** 2
            total += float(w @ flux[at : at + w.size])
            at += w.size
    assert abs(total) < 1e-9
