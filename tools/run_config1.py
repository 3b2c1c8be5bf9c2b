"""BASELINE config 1: 2D Poisson-type problem on a uniform quadtree, p=16 (q=14), L=3, DtN merges — the
reference's own CPU-runnable case (`examples/hp_convergence_2D_problems.py:107-195`, problem 1):
  u_xx + u_yy - cos(5y) u_x + sin(5y) u_y = f on [-1,1]^2, u = 0 on the boundary, u = sin(10 pi x) sin(pi y).

    python tools/run_config1.py                 # CUDA path (needs a GPU)
    python tools/run_config1.py --oracle        # NumPy oracle on the host cores
    python tools/run_config1.py --reference     # the unmodified reference on the NumPy `jax` shim (needs /root/reference)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K, LAM = 5, 10


def soln(x):
    return np.sin(np.pi * LAM * x[..., 0]) * np.sin(np.pi * x[..., 1])


def source(x):
    X, Y = x[..., 0], x[..., 1]
    return (-(np.pi**2) * (1 + LAM**2) * soln(x) - np.pi * LAM * np.cos(np.pi * LAM * X) * np.sin(np.pi * Y) * np.cos(K * Y)
            + np.pi * np.sin(np.pi * LAM * X) * np.cos(np.pi * Y) * np.sin(K * Y))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=16)
    ap.add_argument("--L", type=int, default=3)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    p, q, L = args.p, args.p - 2, args.L
    if args.reference:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "jaxshim"))
        sys.path.insert(0, "/root/reference/src")
        import jax.numpy as jnp
        import jaxhps as api

        arr = jnp.array
    else:
        import jaxhps_b200 as api

        arr = np.asarray
    dom = api.Domain(p=p, q=q, root=api.DiscretizationNode2D(xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0), L=L)
    X = np.asarray(dom.interior_points)
    one = np.ones(X.shape[:2])
    kw = dict(source=arr(source(X)), D_xx_coefficients=arr(one), D_yy_coefficients=arr(one),
              D_x_coefficients=arr(-np.cos(K * X[..., 1])), D_y_coefficients=arr(np.sin(K * X[..., 1])))
    g = np.zeros(np.asarray(dom.boundary_points).shape[0])
    best_b = best_s = 1e30
    for _ in range(args.repeat):
        pb = api.PDEProblem(dom, **kw)
        if args.oracle:
            from oracle import hps_oracle as orc

            t0 = time.perf_counter()
            Y, T, v, h = orc.local_solve_stage_uniform_2D_DtN(pb)
            S, gt = orc.merge_stage_uniform_2D_DtN(T, h, L)
            t1 = time.perf_counter()
            u = orc.down_pass_uniform_2D_DtN(g, S, gt, Y, v)
            t2 = time.perf_counter()
        else:
            if not args.reference:
                import torch

                torch.cuda.synchronize()
            t0 = time.perf_counter()
            api.build_solver(pb) if args.reference else api.build_solver(pb, host_device="cuda")
            if not args.reference:
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            u = np.asarray(api.solve(pb, arr(g)))
            t2 = time.perf_counter()
        best_b, best_s = min(best_b, t1 - t0), min(best_s, t2 - t1)
    exact = soln(X)
    kind = "reference on the NumPy jax shim" if args.reference else ("NumPy oracle" if args.oracle else "CUDA (libhps_b200)")
    print(json.dumps(dict(config="2D uniform DtN (hp_convergence problem 1)", p=p, q=q, L=L, n_leaves=4**L, impl=kind,
                          cores=os.cpu_count(), build_s=round(best_b, 5), solve_s=round(best_s, 5),
                          rel_linf_error=float(np.abs(u - exact).max() / np.abs(exact).max()))))


if __name__ == "__main__":
    main()
