"""BASELINE config 5: linearised Poisson-Boltzmann on an adaptive octree (reference
`examples/poisson_boltzmann_example.py:62-110,226-253`, `examples/poisson_boltzmann_utils.py:13-163`).
  div(eps grad u) = -rho on [-1,1]^3, u = 0 on the boundary,
  rho = sum_i exp(-45 |x - c_i|^2) over 50 atom centres c_i ~ U(-0.5, 0.5)^3,
  eps = 16 + 84 exp(-10 rho);  operator form: eps (u_xx + u_yy + u_zz) + grad(eps).grad(u).
The reference draws the centres with jax.random.key(0), which cannot be reproduced without JAX; they are
drawn with numpy.random.default_rng(0) instead.  The octree is refined on [eps, d_x eps, d_y eps, d_z eps, rho].

    python tools/run_config5.py --p 10 --tol 1e-3 [--mesh-only]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DELTA, EPS_0, EPS_INF, A = 45.0, 16.0, 100.0, 10.0
CENTERS = np.random.default_rng(0).uniform(-0.5, 0.5, size=(50, 3))


def _e(x):
    """(flat points (n,3), exp(-delta |x - c_i|^2) (n,50)); distances through one matrix product."""
    flat = np.asarray(x, dtype=np.float64).reshape(-1, 3)
    d2 = (flat * flat).sum(axis=1)[:, None] - 2.0 * flat @ CENTERS.T + (CENTERS * CENTERS).sum(axis=1)[None]
    return flat, np.exp(-DELTA * np.maximum(d2, 0.0))


def rho(x):
    return _e(x)[1].sum(axis=-1).reshape(np.shape(x)[:-1])


def permittivity(x):
    return EPS_0 + (EPS_INF - EPS_0) * np.exp(-A * rho(x))


def _d_perm(x, axis):
    flat, e = _e(x)
    r = e.sum(axis=-1)
    d_rho = -2 * DELTA * (flat[:, axis] * r - e @ CENTERS[:, axis])  # sum_i -2 delta (x - c_i) e_i
    return (-A * (EPS_INF - EPS_0) * d_rho * np.exp(-A * r)).reshape(np.shape(x)[:-1])


def d_perm_x(x):
    return _d_perm(x, 0)


def d_perm_y(x):
    return _d_perm(x, 1)


def d_perm_z(x):
    return _d_perm(x, 2)


def encode_tree(node):
    """Pre-order list of has-children flags."""
    out = [1 if node.children else 0]
    for c in node.children:
        out += encode_tree(c)
    return out


def decode_tree(root, flags, q):
    from jaxhps_b200._tree import add_eight_children

    it = iter(int(f) for f in flags)

    def walk(node):
        if next(it):
            add_eight_children(node, root=root, q=q)
            for c in node.children:
                walk(c)

    walk(root)
    return root


def build_problem(dom):
    import jaxhps_b200 as hps

    X = dom.interior_points
    eps = permittivity(X)
    return hps.PDEProblem(dom, source=-rho(X), D_xx_coefficients=eps, D_yy_coefficients=eps, D_zz_coefficients=eps,
                          D_x_coefficients=d_perm_x(X), D_y_coefficients=d_perm_y(X), D_z_coefficients=d_perm_z(X))


def main_sharded(args):
    """One process per GPU (torchrun): root octants are split over the ranks (`jaxhps_b200/_dist_adaptive.py`)."""
    import torch
    import torch.distributed as dist

    import jaxhps_b200 as hps
    from jaxhps_b200 import _dist_adaptive as da

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    root = hps.DiscretizationNode3D(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0)
    dom = hps.Domain(p=args.p, q=args.p - 2, root=decode_tree(root, np.load(args.load_tree), args.p - 2))
    pb = build_problem(dom)
    g = dom.get_adaptive_boundary_data_lst(lambda x: np.zeros(x.shape[:-1]))
    ops = da.CudaAdaptiveOps(torch.device("cuda", local))

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t) / 1e3

    for _ in range(args.repeat):
        st, t_build = timed(lambda: da.build_solver_sharded_adaptive(pb, ops))
        (u, sl), t_solve = timed(lambda: da.solve_sharded_adaptive(st, g, ops))
    parts = [torch.empty((n, u.shape[1], u.shape[2]), dtype=torch.float64, device="cuda") for n in st["shard"].leaves_per_rank]
    if world > 1:
        dist.all_gather(parts, u.contiguous())
    else:
        parts = [u]
    if rank == 0:
        u_all = torch.cat(parts)[..., 0].cpu().numpy()
        rec = dict(config="Poisson-Boltzmann adaptive 3D, subtree-sharded", n_gpus=world, p=args.p, q=args.p - 2,
                   n_leaves=dom.n_leaves, leaves_per_rank=st["shard"].leaves_per_rank, build_s=round(t_build, 4),
                   solve_s=round(t_solve, 4), u_max=float(np.abs(u_all).max()),
                   gpu_mem_GB=round(torch.cuda.max_memory_allocated() / 2**30, 2))
        probe_file = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                  f"config5_oracle_probe_p{args.p}.npz")
        if os.path.exists(probe_file):
            ref = np.load(probe_file)
            mine = u_all.reshape(-1)[:: int(ref["stride"])]
            if mine.shape == ref["u_probe"].shape:
                rec["rel_err_vs_oracle"] = float(np.abs(mine - ref["u_probe"]).max() / np.abs(ref["u_probe"]).max())
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=10)
    ap.add_argument("--tol", type=float, default=1e-3)
    ap.add_argument("--mesh-only", action="store_true")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--save-tree", default=None, help="write the refinement pattern (pre-order has-children flags) here")
    ap.add_argument("--load-tree", default=None, help="skip mesh generation and rebuild the octree from this file")
    ap.add_argument("--prof", action="store_true", help="per-category device times of the library's kernels (CUDA events)")
    ap.add_argument("--sharded", action="store_true", help="subtree-sharded build over the ranks of a torchrun launch")
    args = ap.parse_args()
    if args.sharded:
        return main_sharded(args)
    import jaxhps_b200 as hps
    from jaxhps_b200._tree import get_all_leaves

    root = hps.DiscretizationNode3D(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0)
    t0 = time.perf_counter()
    if args.load_tree:
        dom = hps.Domain(p=args.p, q=args.p - 2, root=decode_tree(root, np.load(args.load_tree), args.p - 2))
    else:
        dom = hps.Domain.from_adaptive_discretization(p=args.p, q=args.p - 2, root=root,
                                                      f=[permittivity, d_perm_x, d_perm_y, d_perm_z, rho], tol=args.tol)
    t_mesh = time.perf_counter() - t0
    if args.save_tree:
        np.save(args.save_tree, np.array(encode_tree(root), dtype=np.uint8))
    depths = [leaf.depth for leaf in get_all_leaves(root)]
    rec = dict(config="Poisson-Boltzmann adaptive 3D", p=args.p, q=args.p - 2, tol=args.tol, n_leaves=dom.n_leaves,
               max_depth=max(depths), min_depth=min(depths), n_boundary=int(dom.boundary_points.shape[0]),
               leaves_per_root_octant=[len(get_all_leaves(c)) for c in root.children], mesh_s=round(t_mesh, 2))
    if not args.mesh_only:
        import torch

        pb = build_problem(dom)
        g = dom.get_adaptive_boundary_data_lst(lambda x: np.zeros(x.shape[:-1]))
        import ctypes

        from jaxhps_b200 import _lib

        for rep in range(args.repeat):
            if args.prof and rep == args.repeat - 1:
                _lib.load().hps_prof_enable(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_ev.record()
            hps.build_solver(pb, host_device="cuda")
            b_ev.record()
            torch.cuda.synchronize()
            t_build = time.perf_counter() - t0
            rec["build_device_s"] = round(a_ev.elapsed_time(b_ev) / 1e3, 4)
            if args.prof and rep == args.repeat - 1:
                pm, pw, pl = (ctypes.c_double * 16)(), (ctypes.c_double * 16)(), (ctypes.c_int64 * 16)()
                allk = ctypes.c_int64()
                _lib.load().hps_prof_read(_lib.stream_ptr(), pm, pw, pl, ctypes.byref(allk))
                _lib.load().hps_prof_enable(0)
                names = ["gemm", "lu_panel", "trtri", "laswp", "inner_trsm", "gather", "skinny", "leaf_assemble"]
                rec["prof_ms"] = {n: round(pm[i], 1) for i, n in enumerate(names)}
                rec["prof_launches"] = {n: int(pl[i]) for i, n in enumerate(names)}
                rec["gemm_TFLOPs"] = round(pw[0] / 1e12, 2)
                rec["all_launches"] = int(allk.value)
            t0 = time.perf_counter()
            u = hps.solve(pb, g)
            t_solve = time.perf_counter() - t0
        # repeated solves with the same build: the second one captures the down pass as a CUDA graph, later ones replay it
        t_rep = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            u2 = hps.solve(pb, g, host_device="cuda")
            torch.cuda.synchronize()
            t_rep.append(time.perf_counter() - t0)
        rec["solve_repeat_s"] = [round(t, 5) for t in t_rep]
        rec["graph_replay_max_diff"] = float((u2.cpu() - torch.as_tensor(u)).abs().max())
        probe_file = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                  f"config5_oracle_probe_p{args.p}.npz")
        if os.path.exists(probe_file) and args.load_tree:
            ref = np.load(probe_file)
            if ref["u_probe"].shape == u.reshape(-1)[:: int(ref["stride"])].shape:
                rec["rel_err_vs_oracle"] = float(np.abs(u.reshape(-1)[:: int(ref["stride"])] - ref["u_probe"]).max()
                                                 / np.abs(ref["u_probe"]).max())
        rec.update(build_s=round(t_build, 4), solve_s=round(t_solve, 4), u_max=float(np.abs(u).max()),
                   gpu_mem_GB=round(torch.cuda.max_memory_allocated() / 2**30, 2))
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
