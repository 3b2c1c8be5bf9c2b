"""Adaptive host layer (trees, mesh generation, point clouds, projection operators) against golden
vectors produced by the unmodified reference (tests/golden/make_golden_adaptive.py)."""
import os
import sys

import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200 import _operators as ops
from jaxhps_b200._tree import add_eight_children, add_four_children, get_all_leaves

from _cases import GOLDEN_DIR

sys.path.insert(0, GOLDEN_DIR)
from adaptive_cases import ADAPTIVE_CASES, build_domain, internal_nodes  # noqa: E402


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def our_domain(case):
    add = add_four_children if case["dim"] == 2 else add_eight_children
    return build_domain(hps, lambda n, r, q: add(n, root=r, q=q), case)


@pytest.mark.parametrize("name", sorted(ADAPTIVE_CASES))
def test_tree_and_point_clouds_match_reference(name):
    case, gold = ADAPTIVE_CASES[name], load(name)
    dom = our_domain(case)
    nf = 2 * case["dim"]
    keys = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[:nf]
    bounds = np.array([[getattr(l, k) for k in keys] for l in get_all_leaves(dom.root)])
    assert bounds.shape == gold["leaf_bounds"].shape and np.array_equal(bounds, gold["leaf_bounds"])
    assert not dom.bool_uniform and dom.n_leaves == bounds.shape[0]
    np.testing.assert_allclose(dom.interior_points, gold["interior_points"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(dom.boundary_points, gold["boundary_points"], rtol=0, atol=1e-15)
    n_sides = np.array([[getattr(n, f"n_{f}") for f in range(nf)] for n in internal_nodes(dom.root)])
    assert np.array_equal(n_sides, gold["n_sides"])


def test_projection_and_refinement_operators_match_reference():
    gold = load("adapt_operators")
    for q in (2, 4, 6):
        a, b = ops.precompute_projection_ops_3D(q)
        np.testing.assert_allclose(a, gold[f"L_4f1_q{q}"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(b, gold[f"L_1f4_q{q}"], rtol=0, atol=1e-14)
        a, b = ops.precompute_projection_ops_2D(q)
        np.testing.assert_allclose(a, gold[f"L_2f1_q{q}"], rtol=0, atol=1e-14)
        np.testing.assert_allclose(b, gold[f"L_1f2_q{q}"], rtol=0, atol=1e-14)
    for p in (4, 5):
        np.testing.assert_allclose(ops.precompute_L_4f1(p), gold[f"L_4f1_cheb_p{p}"], rtol=0, atol=1e-13)
    for p in (3, 4):
        np.testing.assert_allclose(ops.precompute_L_8f1(p), gold[f"L_8f1_cheb_p{p}"], rtol=0, atol=1e-13)
    with pytest.raises(ValueError):
        ops.precompute_projection_ops_3D(3)


def test_boundary_data_list_and_counts():
    case = ADAPTIVE_CASES["adapt3d_p4q2_manual"]
    dom = our_domain(case)
    lst = dom.get_adaptive_boundary_data_lst(lambda x: x[..., 0])
    assert [g.shape[0] for g in lst] == [getattr(dom.root, f"n_{f}") for f in range(6)]
    assert sum(g.shape[0] for g in lst) == dom.boundary_points.shape[0]
    # splitting twice is a no-op (the reference guards the counts the same way)
    before = dom.root.n_1
    add_eight_children(dom.root.children[2], root=dom.root, q=case["q"])
    assert dom.root.n_1 == before
