"""Shared, seeded problem builders for the tests (same recipe as tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

import jaxhps_b200 as hps

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "ref*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def make_domain(dim, p, q, L):
    if dim == 3:
        root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    else:
        root = hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0)
    return hps.Domain(p, q, root, L)


def seeded_inputs(dim, p, q, L, nsrc, seed):
    rng = np.random.default_rng(seed)
    n_leaves = (8 if dim == 3 else 4) ** L
    shp = (n_leaves, p**dim)
    names = ["D_xx", "D_yy"] + (["D_zz"] if dim == 3 else [])
    co = {f"{k}_coefficients": 1 + 0.1 * rng.normal(size=shp) for k in names}
    co["D_xy_coefficients"] = 0.1 * rng.normal(size=shp)
    co["D_y_coefficients"] = rng.normal(size=shp)
    co["I_coefficients"] = rng.normal(size=shp)
    if dim == 3:
        co["D_z_coefficients"] = rng.normal(size=shp)
        co["D_yz_coefficients"] = 0.1 * rng.normal(size=shp)
    src = rng.normal(size=shp if nsrc == 1 else shp + (nsrc,))
    n_bdry = (6 * 4**L * q * q) if dim == 3 else (4 * 2**L * q)
    bdry = rng.normal(size=(n_bdry,) if nsrc == 1 else (n_bdry, nsrc))
    return co, src, bdry


ETA = 4.0  # impedance parameter of the ItI cases (dim code 20 = 2D ItI)


def seeded_inputs_iti(p, q, L, nsrc, seed):
    rng = np.random.default_rng(seed)
    shp = (4**L, p * p)
    co = {"D_xx_coefficients": np.ones(shp), "D_yy_coefficients": np.ones(shp),
          "I_coefficients": ETA**2 * (1 + 0.3 * rng.normal(size=shp))}
    sshape = shp if nsrc == 1 else shp + (nsrc,)
    src = rng.normal(size=sshape) + 1j * rng.normal(size=sshape)
    n_bdry = 4 * 2**L * q
    bshape = (n_bdry,) if nsrc == 1 else (n_bdry, nsrc)
    bdry = rng.normal(size=bshape) + 1j * rng.normal(size=bshape)
    return co, src, bdry


def seeded_problem(dim, p, q, L, nsrc=1, seed=0):
    if dim == 20:
        co, src, bdry = seeded_inputs_iti(p, q, L, nsrc, seed)
        dom = make_domain(2, p, q, L)
        return hps.PDEProblem(dom, source=src, use_ItI=True, eta=ETA, **co), bdry
    co, src, bdry = seeded_inputs(dim, p, q, L, nsrc, seed)
    dom = make_domain(dim, p, q, L)
    return hps.PDEProblem(dom, source=src, **co), bdry


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ---- BASELINE config 3 at full leaf size (p=12, q=10): parity problem and the probe fixtures ----
CONFIG3_P, CONFIG3_Q, CONFIG3_SEED = 12, 10, 3


def config3_problem(L):
    """SURVEY §8(d) config 3, random-coefficient variant for parity: D_xx, D_yy, D_zz = 1 + 0.1 N(0,1) (seed 3),
    a first-order and a zeroth-order term so that every merge sees a non-symmetric operator, random source and
    random Dirichlet data."""
    rng = np.random.default_rng(CONFIG3_SEED)
    shp = (8**L, CONFIG3_P**3)
    co = {f"{k}_coefficients": 1 + 0.1 * rng.normal(size=shp) for k in ("D_xx", "D_yy", "D_zz")}
    co["D_x_coefficients"] = rng.normal(size=shp)
    co["I_coefficients"] = rng.normal(size=shp)
    src = rng.normal(size=shp)
    bdry = rng.normal(size=6 * 4**L * CONFIG3_Q**2)
    dom = make_domain(3, CONFIG3_P, CONFIG3_Q, L)
    return hps.PDEProblem(dom, source=src, **co), bdry


def config3_probe(n, salt):
    """Fixed probe vector of length n (operators are compared through their action on it)."""
    return np.random.default_rng(1000 + salt).normal(size=n)


# ---- scattering coupling: seeded inputs shared by the golden generator and the tests ----
def scattering_inputs(n_side=12, n_src=3, seed=7):
    """A random complex ItI-like matrix R (spectrum away from 1), synthetic single / double layer matrices S, D
    (the reference loads them from MATLAB files), boundary points of [-1,1]^2 in the S, E, N, W order and a few
    plane-wave directions."""
    rng = np.random.default_rng(seed)
    n = 4 * n_side
    R = 0.3 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n)
    S = 0.2 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n)
    D = 0.2 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n)
    t = (np.arange(n_side) + 0.5) / n_side * 2 - 1
    one = np.ones(n_side)
    pts = np.concatenate([np.stack([t, -one], 1), np.stack([one, t], 1), np.stack([-t, one], 1), np.stack([-one, -t], 1)])
    dirs = rng.uniform(0, 2 * np.pi, size=n_src)
    return R, S, D, pts, dirs, 5.0, 5.0  # k, eta


# ---- BASELINE config 2: 2D variable-coefficient Helmholtz, ItI, p=16 q=14, k=100 ----
def config2_problem(L, k=100.0, half_width=1.0):
    """``u_xx + u_yy + k^2 (1 + q(x)) u = -k^2 q e^{ikx}`` on [-1,1]^2, q = sum of 10 gauss bumps (centres U(-0.5,0.5)^2,
    seed 0), eta = k; boundary data = incoming impedance of the plane wave e^{ikx}.  ``half_width`` shrinks the domain
    (bumps scaled with it): ``half_width = 2**L / 64`` keeps config 2's leaf size 1/32 — hence its leaf conditioning —
    on a tree shallow enough for the extended-precision arbitration."""
    w = float(half_width)
    dom = hps.Domain(16, 14, hps.DiscretizationNode2D(-w, w, -w, w), L)
    x = dom.interior_points
    one = np.ones_like(x[..., 0])
    rng = np.random.default_rng(0)
    centres = rng.uniform(-0.5, 0.5, size=(10, 2)) * w
    q = sum(np.exp(-50 * ((x[..., 0] - c[0]) ** 2 + (x[..., 1] - c[1]) ** 2) / w**2) for c in centres)
    pb = hps.PDEProblem(dom, source=-k**2 * q * np.exp(1j * k * x[..., 0]), D_xx_coefficients=one, D_yy_coefficients=one,
                        I_coefficients=k**2 * (1 + q), use_ItI=True, eta=k)
    b = dom.boundary_points
    n = b.shape[0] // 4
    ub = np.exp(1j * k * b[:, 0])
    nx = np.concatenate([np.zeros(n), np.ones(n), np.zeros(n), -np.ones(n)])
    return dom, pb, nx * 1j * k * ub + 1j * k * ub
