"""BASELINE config 4: 3D wavefront problem on an adaptive octree (reference
`examples/wavefront_adaptive_discretization_3D.py:297-400`, `examples/wavefront_data.py:11-233`).
u = arctan(10 (r - 0.7)), r = |x + 0.05|, on [0,1]^3; source = Laplacian(u); Dirichlet data = u.

    python tools/run_config4.py --p 10 --tol 1e-2 1e-3 [--mesh-only]
"""
import argparse
import json
import time

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def radius(x):
    return np.sqrt((x[..., 0] + 0.05) ** 2 + (x[..., 1] + 0.05) ** 2 + (x[..., 2] + 0.05) ** 2)


def wavefront_soln(x):
    return np.arctan(10 * (radius(x) - 0.7))


def source(x):
    r = radius(x)
    s = r - 0.7
    den = 1 + 100 * s * s
    return -2000 * s / den**2 + 2 * 10 / (den * r)  # f'' + 2 f'/r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=10)
    ap.add_argument("--tol", type=float, nargs="+", default=[1e-2])
    ap.add_argument("--mesh-only", action="store_true")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--mesh-device", default=None, help="run the refinement check of the mesh generator there (e.g. cuda:0)")
    args = ap.parse_args()
    import jaxhps_b200 as hps
    from jaxhps_b200._tree import get_all_leaves

    for tol in args.tol:
        root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        t0 = time.perf_counter()
        dom = hps.Domain.from_adaptive_discretization(p=args.p, q=args.p - 2, root=root, f=source, tol=tol, device=args.mesh_device)
        t_mesh = time.perf_counter() - t0
        depths = [leaf.depth for leaf in get_all_leaves(root)]
        rec = dict(config="wavefront adaptive 3D", p=args.p, q=args.p - 2, tol=tol, n_leaves=dom.n_leaves,
                   max_depth=max(depths), min_depth=min(depths), n_boundary=int(dom.boundary_points.shape[0]),
                   mesh_s=round(t_mesh, 3))
        if not args.mesh_only:
            import torch

            one = np.ones(dom.interior_points.shape[:2])
            pb = hps.PDEProblem(dom, source=source(dom.interior_points), D_xx_coefficients=one, D_yy_coefficients=one,
                                D_zz_coefficients=one)
            g = dom.get_adaptive_boundary_data_lst(wavefront_soln)
            for _ in range(args.repeat):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                hps.build_solver(pb, host_device="cuda")
                torch.cuda.synchronize()
                t_build = time.perf_counter() - t0
                t0 = time.perf_counter()
                u = hps.solve(pb, g)
                t_solve = time.perf_counter() - t0
            exact = wavefront_soln(dom.interior_points)
            rec.update(build_s=round(t_build, 4), solve_s=round(t_solve, 4),
                       rel_linf_error=float(np.abs(u - exact).max() / np.abs(exact).max()))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
