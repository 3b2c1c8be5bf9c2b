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
