"""Generate golden vectors by executing the UNMODIFIED reference (/root/reference/src/jaxhps)
on the NumPy-backed `jax` shim in tests/golden/jaxshim (JAX itself is not installable here).

    python tests/golden/make_golden.py        # needs /root/reference; writes tests/golden/*.npz

The fixtures travel with the repo; nothing at test time reads /root/reference.
Inputs are seeded with numpy.random.default_rng so the tests can rebuild them exactly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference/src")

import jax.numpy as jnp  # noqa: E402  (the shim)
import jaxhps as ref  # noqa: E402
from jaxhps.local_solve import (  # noqa: E402
    local_solve_stage_uniform_2D_DtN,
    local_solve_stage_uniform_2D_ItI,
    local_solve_stage_uniform_3D_DtN,
)
from jaxhps.merge import (  # noqa: E402
    merge_stage_uniform_2D_DtN,
    merge_stage_uniform_2D_ItI,
    merge_stage_uniform_3D_DtN,
)
from jaxhps.down_pass import (  # noqa: E402
    down_pass_uniform_2D_DtN,
    down_pass_uniform_2D_ItI,
    down_pass_uniform_3D_DtN,
)

ETA = 4.0  # impedance parameter of the ItI cases (dim code 20 = 2D ItI)


def inputs_iti(p, q, L, nsrc, seed):
    """Seeded Helmholtz-type ItI problem: u_xx + u_yy + k^2 (1 + 0.3 N(0,1)) u = f, complex f and data."""
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


def inputs(dim, p, q, L, nsrc, seed):
    """Seeded variable-coefficient problem; mirrored by tests/_cases.py."""
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


def run_case(dim, p, q, L, nsrc, seed, full):
    iti = dim == 20
    co, src, bdry = inputs_iti(p, q, L, nsrc, seed) if iti else inputs(dim, p, q, L, nsrc, seed)
    if iti:
        root = ref.DiscretizationNode2D(xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0)
        ls, mg, dp = local_solve_stage_uniform_2D_ItI, merge_stage_uniform_2D_ItI, down_pass_uniform_2D_ItI
    elif dim == 3:
        root = ref.DiscretizationNode3D(xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, zmin=0.0, zmax=1.0)
        ls, mg, dp = local_solve_stage_uniform_3D_DtN, merge_stage_uniform_3D_DtN, down_pass_uniform_3D_DtN
    else:
        root = ref.DiscretizationNode2D(xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0)
        ls, mg, dp = local_solve_stage_uniform_2D_DtN, merge_stage_uniform_2D_DtN, down_pass_uniform_2D_DtN
    dom = ref.Domain(p=p, q=q, root=root, L=L)
    extra = dict(use_ItI=True, eta=ETA) if iti else {}
    pb = ref.PDEProblem(dom, source=jnp.array(src), **{k: jnp.array(v) for k, v in co.items()}, **extra)
    Y, T, v, h = ls(pb)
    S_lst, g_lst, T_top = mg(T, h, l=L, return_T=True)
    u = dp(jnp.array(bdry), S_lst, g_lst, Y, v)
    # T_top is large; every case stores its action on a seeded probe, full cases store it whole
    probe = np.random.default_rng(seed + 1000).normal(size=np.asarray(T_top).shape[1])
    out = dict(meta=np.array([dim, p, q, L, nsrc, seed]), u=np.asarray(u), T_top_probe=np.asarray(T_top) @ probe,
               v=np.asarray(v), h=np.asarray(h))
    for i, g in enumerate(g_lst):
        out[f"g_tilde_{i}"] = np.asarray(g)
    if full:
        out.update(Y=np.asarray(Y), T=np.asarray(T), T_top=np.asarray(T_top), P=np.asarray(pb.P),
                   D_x=np.asarray(pb.D_x), interior_points=np.asarray(dom.interior_points),
                   boundary_points=np.asarray(dom.boundary_points))
        if iti:
            out.update(G=np.asarray(pb.G), QH=np.asarray(pb.QH))
        else:
            out.update(Q=np.asarray(pb.Q))
        for i, S in enumerate(S_lst):
            out[f"S_{i}"] = np.asarray(S)
    return out


CASES = {
    # name: (dim, p, q, L, nsrc, seed, store-everything?)
    "ref3d_p4q2L1": (3, 4, 2, 1, 1, 11, True),
    "ref3d_p4q2L2": (3, 4, 2, 2, 1, 12, False),
    # 3D multi-source needs L >= 2 in the reference (down_pass/_uniform_3D_DtN.py:70)
    "ref3d_p4q2L2_ms": (3, 4, 2, 2, 2, 13, False),
    "ref3d_p5q3L1": (3, 5, 3, 1, 1, 14, True),
    "ref2d_p6q4L2": (2, 6, 4, 2, 1, 21, True),
    "ref2d_p7q5L1_ms": (2, 7, 5, 1, 3, 22, True),
    "refiti_p6q4L2": (20, 6, 4, 2, 1, 31, True),
    "refiti_p8q6L2_ms": (20, 8, 6, 2, 2, 32, False),
}

def interp_case():
    """Domain.interp_{to,from}_interior_points of the reference on seeded data (2D and 3D)."""
    rng = np.random.default_rng(77)
    out = {}
    d2 = ref.Domain(p=6, q=4, root=ref.DiscretizationNode2D(xmin=-1.0, xmax=1.0, ymin=0.0, ymax=2.0), L=2)
    f2 = rng.normal(size=(16, 36))
    xs, ys = np.linspace(-1, 1, 7), np.linspace(0, 2, 5)
    v, t = d2.interp_from_interior_points(jnp.array(f2), jnp.array(xs), jnp.array(ys))
    g2 = rng.normal(size=(9, 8))
    sx, sy = np.linspace(-1, 1, 9), np.linspace(0, 2, 8)
    out.update(f2=f2, from2=np.asarray(v), pts2=np.asarray(t), g2=g2,
               to2=np.asarray(d2.interp_to_interior_points(jnp.array(g2), jnp.array(sx), jnp.array(sy))))
    d3 = ref.Domain(p=4, q=2, root=ref.DiscretizationNode3D(xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, zmin=-1.0, zmax=0.0), L=1)
    f3 = rng.normal(size=(8, 64))
    x3, y3, z3 = np.linspace(0, 1, 5), np.linspace(0, 1, 4), np.linspace(-1, 0, 3)
    v, t = d3.interp_from_interior_points(jnp.array(f3), jnp.array(x3), jnp.array(y3), jnp.array(z3))
    g3 = rng.normal(size=(5, 6, 7))
    out.update(f3=f3, from3=np.asarray(v), pts3=np.asarray(t), g3=g3,
               to3=np.asarray(d3.interp_to_interior_points(jnp.array(g3), jnp.array(np.linspace(0, 1, 5)),
                                                           jnp.array(np.linspace(0, 1, 6)), jnp.array(np.linspace(-1, 0, 7)))))
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "interp_reference.npz"), **interp_case())
    for name, args in CASES.items():
        data = run_case(*args)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, {k: v.shape for k, v in data.items() if k != "meta"}, os.path.getsize(path) // 1024, "KiB")
