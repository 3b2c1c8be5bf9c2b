"""GPU parity of the interpolation kernels against the reference-generated fixture
(tests/golden/interp_reference.npz) and against the host implementation on larger seeded cases incl. adaptive trees
and multi-source data; 1e-10 relative."""
import os

import numpy as np
import pytest

import jaxhps_b200 as hps
from _cases import GOLDEN_DIR, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10
DEV = "cuda:0"


def test_device_interpolation_matches_reference_fixture():
    G = dict(np.load(os.path.join(GOLDEN_DIR, "interp_reference.npz")))
    d2 = hps.Domain(6, 4, hps.DiscretizationNode2D(-1.0, 1.0, 0.0, 2.0), 2)
    v, t = d2.interp_from_interior_points(G["f2"], np.linspace(-1, 1, 7), np.linspace(0, 2, 5), device=DEV, host_device="cpu")
    assert rel_err(v, G["from2"]) < TOL and np.array_equal(t, G["pts2"])
    to = d2.interp_to_interior_points(G["g2"], np.linspace(-1, 1, 9), np.linspace(0, 2, 8), device=DEV, host_device="cpu")
    assert rel_err(to, G["to2"]) < TOL
    d3 = hps.Domain(4, 2, hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, -1.0, 0.0), 1)
    v, t = d3.interp_from_interior_points(G["f3"], np.linspace(0, 1, 5), np.linspace(0, 1, 4), np.linspace(-1, 0, 3),
                                          device=DEV, host_device="cpu")
    assert rel_err(v, G["from3"]) < TOL and np.array_equal(t, G["pts3"])
    to = d3.interp_to_interior_points(G["g3"], np.linspace(0, 1, 5), np.linspace(0, 1, 6), np.linspace(-1, 0, 7),
                                      device=DEV, host_device="cpu")
    assert rel_err(to, G["to3"]) < TOL


@pytest.mark.parametrize("dim,p,L", [(2, 16, 3), (3, 12, 2), (3, 7, 1)])
def test_device_interpolation_matches_host(dim, p, L):
    rng = np.random.default_rng(p + L)
    if dim == 2:
        dom = hps.Domain(p, p - 2, hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0), L)
        grids = (np.linspace(-1, 1, 41), np.linspace(-1, 1, 37))
        from_grids = (np.cos(np.pi * np.arange(14) / 13)[::-1], np.linspace(-1, 1, 11))
    else:
        dom = hps.Domain(p, p - 2, hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0), L)
        grids = (np.linspace(0, 1, 17), np.linspace(0, 1, 13), np.linspace(0, 1, 11))
        from_grids = (np.linspace(0, 1, 9), np.linspace(0, 1, 8), 0.5 + 0.5 * np.cos(np.pi * np.arange(10) / 9)[::-1])
    f = rng.normal(size=dom.interior_points[..., 0].shape)
    v_h, t_h = dom.interp_from_interior_points(f, *grids)
    v_d, t_d = dom.interp_from_interior_points(f, *grids, device=DEV, host_device="cpu")
    assert rel_err(v_d, v_h) < TOL and np.array_equal(t_d, t_h)
    if dim == 2:  # trailing source axis (2D only, like the reference)
        fm = rng.normal(size=f.shape + (3,))
        vm_h, _ = dom.interp_from_interior_points(fm, *grids)
        vm_d, _ = dom.interp_from_interior_points(fm, *grids, device=DEV, host_device="cpu")
        assert rel_err(vm_d, vm_h) < TOL
    vals = rng.normal(size=tuple(g.shape[0] for g in from_grids))
    to_h = dom.interp_to_interior_points(vals, *from_grids)
    to_d = dom.interp_to_interior_points(vals, *from_grids, device=DEV, host_device="cpu")
    assert rel_err(to_d, to_h) < TOL


def test_device_interpolation_on_an_adaptive_tree():
    root = hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0)
    dom = hps.Domain.from_adaptive_discretization(p=8, q=6, root=root, f=lambda x: np.exp(-40 * ((x[..., 0] - 0.3) ** 2 + x[..., 1] ** 2)),
                                                  tol=1e-4)
    x = dom.interior_points
    f = np.sin(3 * x[..., 0]) * np.cos(2 * x[..., 1])
    xs, ys = np.linspace(-0.97, 0.99, 53), np.linspace(-1, 1, 47)
    v_h, _ = dom.interp_from_interior_points(f, xs, ys)
    v_d, _ = dom.interp_from_interior_points(f, xs, ys, device=DEV, host_device="cpu")
    assert rel_err(v_d, v_h) < TOL
