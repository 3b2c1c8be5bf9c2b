"""The scattering-coupling oracle against the reference's own functions (run on the jax shim, frozen in
tests/golden/scattering_reference.npz by tests/golden/make_golden_scattering.py)."""
import os

import numpy as np

from oracle import hps_oracle_scattering as osc
from _cases import GOLDEN_DIR, rel_err, scattering_inputs


def test_oracle_matches_reference_fixture():
    G = dict(np.load(os.path.join(GOLDEN_DIR, "scattering_reference.npz")))
    R, S, D, pts, dirs, k, eta = scattering_inputs()
    T = osc.get_DtN_from_ItI(R, eta)
    assert rel_err(T, G["T"]) < 1e-13
    uin, normals = osc.get_uin_and_normals(k, pts, dirs)
    assert rel_err(uin, G["uin"]) < 1e-14 and rel_err(normals, G["normals"]) < 1e-14
    A, b = osc.setup_scattering_lin_system(S, D, T, pts, k, dirs)
    assert rel_err(A, G["A"]) < 1e-13 and rel_err(b, G["b"]) < 1e-13
    assert rel_err(osc.get_scattering_uscat_impedance(S, D, T, dirs, pts, k, eta), G["imp"]) < 1e-12


def test_host_plane_wave_data_matches_reference_fixture():
    from jaxhps_b200.scattering import get_uin_and_normals

    G = dict(np.load(os.path.join(GOLDEN_DIR, "scattering_reference.npz")))
    _, _, _, pts, dirs, k, _ = scattering_inputs()
    uin, normals = get_uin_and_normals(k, pts, dirs)
    assert rel_err(uin, G["uin"]) < 1e-14 and rel_err(normals, G["normals"]) < 1e-14
