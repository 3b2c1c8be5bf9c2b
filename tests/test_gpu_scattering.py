"""GPU parity of the ItI -> DtN conversion and the BIE coupling solve against the oracle and the
reference-generated fixture, 1e-10; and the same on the top-level ItI operator of a built solver."""
import os

import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200 import scattering as sc
from oracle import hps_oracle_scattering as osc
from _cases import GOLDEN_DIR, rel_err, scattering_inputs, seeded_problem

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_coupling_matches_reference_fixture():
    G = dict(np.load(os.path.join(GOLDEN_DIR, "scattering_reference.npz")))
    R, S, D, pts, dirs, k, eta = scattering_inputs()
    T = sc.get_DtN_from_ItI(R, eta)
    assert rel_err(T, G["T"]) < TOL
    A, b = sc.setup_scattering_lin_system(S, D, T, pts, k, dirs)
    assert rel_err(A, G["A"]) < TOL and rel_err(b, G["b"]) < TOL
    assert rel_err(sc.get_scattering_uscat_impedance(S, D, T, dirs, pts, k, eta), G["imp"]) < TOL


@pytest.mark.parametrize("p,q,L", [(8, 6, 2), (16, 14, 2)])
def test_top_level_iti_operator_to_dtn(p, q, L):
    """R_top of a built ItI solver (return_top_T) -> DtN on the device vs the oracle's conversion of the same R."""
    pb, _ = seeded_problem(20, p, q, L, 1, seed=31)
    R_top = hps.build_solver(pb, return_top_T=True)
    T = sc.get_DtN_from_ItI(R_top, pb.eta)
    To = osc.get_DtN_from_ItI(np.asarray(R_top), pb.eta)
    assert rel_err(T, To) < TOL
    # coupling with synthetic layer potentials on the solver's own boundary points
    n = T.shape[0]
    rng = np.random.default_rng(5)
    S = 0.2 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n)
    D = 0.2 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n)
    dirs = np.array([0.3, 2.0])
    pts = pb.domain.boundary_points
    imp = sc.get_scattering_uscat_impedance(S, D, T, dirs, pts, 3.0, pb.eta)
    assert rel_err(imp, osc.get_scattering_uscat_impedance(S, D, To, dirs, pts, 3.0, pb.eta)) < 1e-10
