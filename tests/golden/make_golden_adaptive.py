"""Golden vectors of the adaptive path: the UNMODIFIED reference executed on the NumPy `jax` shim.

    python tests/golden/make_golden_adaptive.py     # needs /root/reference; writes tests/golden/adapt_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, HERE)

import jax.numpy as jnp  # noqa: E402  (the shim)
import jaxhps as ref  # noqa: E402
from jaxhps._discretization_tree import get_all_leaves  # noqa: E402
from jaxhps._discretization_tree_operations_2D import add_four_children  # noqa: E402
from jaxhps._discretization_tree_operations_3D import add_eight_children  # noqa: E402
from jaxhps._precompute_operators_2D import precompute_L_4f1, precompute_projection_ops_2D  # noqa: E402
from jaxhps._precompute_operators_3D import precompute_L_8f1, precompute_projection_ops_3D  # noqa: E402

from adaptive_cases import ADAPTIVE_CASES, boundary_fn, build_domain, internal_nodes, seeded_fields  # noqa: E402


def run(case):
    dim, p, q = case["dim"], case["p"], case["q"]
    add = (lambda n, r, q_: add_four_children(n, root=r, q=q_)) if dim == 2 else (lambda n, r, q_: add_eight_children(n, root=r, q=q_))
    dom = build_domain(ref, add, case, xp=jnp)
    root = dom.root
    leaves = get_all_leaves(root)
    nf = 2 * dim
    bounds = np.array([[float(getattr(l, k)) for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[:nf]] for l in leaves])
    out = dict(leaf_bounds=bounds, interior_points=np.asarray(dom.interior_points), boundary_points=np.asarray(dom.boundary_points))
    nodes = internal_nodes(root)
    out["n_sides"] = np.array([[int(getattr(n, f"n_{f}")) for f in range(nf)] for n in nodes])
    co, src = seeded_fields(case, len(leaves))
    pb = ref.PDEProblem(dom, source=jnp.array(src), **{k: jnp.array(v) for k, v in co.items()})
    T_top = np.asarray(ref.build_solver(pb, return_top_T=True))
    g_lst = dom.get_adaptive_boundary_data_lst(lambda x: boundary_fn(x, jnp))
    u = np.asarray(ref.solve(pb, g_lst))
    rng = np.random.default_rng(case["seed"] + 1000)
    out.update(u=u, v=np.stack([np.asarray(l.data.v) for l in leaves]), T_top_probe=T_top @ rng.normal(size=T_top.shape[1]),
               g_bdry=np.concatenate([np.asarray(g) for g in g_lst]))
    small = T_top.shape[0] <= 400
    if small:
        out.update(T_top=T_top, Y=np.stack([np.asarray(l.data.Y) for l in leaves]), T_leaf=np.stack([np.asarray(l.data.T) for l in leaves]),
                   h_leaf=np.stack([np.asarray(l.data.h) for l in leaves]))
    for i, n in enumerate(nodes):
        S = np.asarray(n.data.S)
        out[f"g_tilde_{i}"] = np.asarray(n.data.g_tilde)
        out[f"h_{i}"] = np.asarray(n.data.h)
        out[f"S_probe_{i}"] = S @ rng.normal(size=S.shape[1])
        if small:
            out[f"S_{i}"] = S
    return out


def operators():
    out = {}
    for q in (2, 4, 6):
        a, b = precompute_projection_ops_3D(q)
        out[f"L_4f1_q{q}"], out[f"L_1f4_q{q}"] = np.asarray(a), np.asarray(b)
        a, b = precompute_projection_ops_2D(q)
        out[f"L_2f1_q{q}"], out[f"L_1f2_q{q}"] = np.asarray(a), np.asarray(b)
    for p in (4, 5):
        out[f"L_4f1_cheb_p{p}"] = np.asarray(precompute_L_4f1(p))
    for p in (3, 4):
        out[f"L_8f1_cheb_p{p}"] = np.asarray(precompute_L_8f1(p))
    return out


if __name__ == "__main__":
    if len(sys.argv) == 1:
        np.savez_compressed(os.path.join(HERE, "adapt_operators.npz"), **operators())
    only = set(sys.argv[1:])
    for name, case in ADAPTIVE_CASES.items():
        if only and name not in only:
            continue
        data = run(case)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, "leaves", data["leaf_bounds"].shape[0], "n_bdry", data["g_bdry"].shape[0], os.path.getsize(path) // 1024, "KiB")
