#!/usr/bin/env python
"""bench.py — HPS build_solver + solve on B200, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--L L]

One "step" = one full pass of the hot path over the batch of leaves:
local_solve_stage -> merge_stage (all levels) -> down_pass, 3D Poisson-type operator with a
synthetic variable coefficient field, p=12, q=10, FP64 (BASELINE config 3).  At N=1 the tree
has L=3 levels (512 leaves) — the largest configuration whose operators fit one GPU (L=4 needs
a 76 800^2 root merge, 47 GB for D alone plus 94 GB for S).

value      = leaves per second through build+solve with inputs resident in HBM.
e2e        = same through the public API with HOST (pinned) inputs and the solution read back.
roofline   = DMMA GEMM kernel: algorithmic flops / summed CUDA-event launch time vs the measured
             cuBLAS DGEMM rate on this pool's B200s (profiles/r01_fp64_probe.txt).
cpu_baseline / --impl reference = the NumPy oracle (restatement of the reference's algorithm;
             the reference's JAX runtime is not installable) on the box's host cores, on a
             bounded sample of the workload (one depth-1 subtree: 8 leaves, 1 merge, 1 solve).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, Q = 12, 10
# measured on this pool's B200 (profiles/r01_fp64_probe.txt): cuBLAS DGEMM 8192^3
FP64_PEAK_TFLOPS = 35.4
FP64_PEAK_TFLOPS_SUSTAINED = 35.4


def u_exact(x):
    """Manufactured solution used by every bench problem: u = sin(2 pi x) cos(pi y) exp(z)."""
    return np.sin(2 * np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1]) * np.exp(x[..., 2])


def synthetic_fields(L, p=P, leaf_slice=None):
    """Synthetic variable-coefficient problem with a known answer:
    c(x) (u_xx + u_yy + u_zz) = f on [0,1]^3 with c = 1 + 0.5 exp(-|x-1/2|^2/0.1) on D_xx, D_yy, D_zz,
    f = c (1 - 5 pi^2) u_exact and Dirichlet data u_exact on the boundary, so each run can report
    its own error against the analytic solution."""
    import jaxhps_b200 as hps

    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    dom = hps.Domain(p, Q, root, L)
    x = dom.interior_points
    r2 = ((x - 0.5) ** 2).sum(axis=-1)
    c = 1.0 + 0.5 * np.exp(-r2 / 0.1)
    src = c * (1.0 - 5.0 * np.pi**2) * u_exact(x)
    g = u_exact(dom.boundary_points)
    return dom, c, src, g


def lean_flops(L, p=P, q=Q, root_T=False):
    """Algorithmic flops of one build: LU+solve formulation, block-sparse B."""
    n_i, n_b, n_c, n_g = (p - 2) ** 3, p**3 - (p - 2) ** 3, p**3, 6 * q * q
    leaf = (2 / 3) * n_i**3 + 2 * n_i**2 * (n_g + 1) + 2 * n_i * n_b * n_g + 2 * n_g * n_c * n_g
    total = 8**L * leaf
    m = q * q
    for level in range(L, 0, -1):
        n_merges = 8 ** (level - 1)
        per = 11520 if (level > 1 or root_T) else 8064
        total += n_merges * per * float(m) ** 3
        m *= 4
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- CPU arm


def cpu_sample_step(pb_sample, g_sample):
    from oracle import hps_oracle as orc

    Y, T, v, h = orc.local_solve_stage_uniform_3D_DtN(pb_sample)
    S, gt = orc.merge_stage_uniform_3D_DtN(T, h, 1)
    return orc.down_pass_uniform_3D_DtN(g_sample, S, gt, Y, v)


def cpu_sample_problem():
    """One depth-1 subtree of the workload: 8 leaves of the p=12 problem, one oct merge, one solve."""
    import jaxhps_b200 as hps

    dom, c, src, g = synthetic_fields(1)
    pb = hps.PDEProblem(dom, source=src, D_xx_coefficients=c, D_yy_coefficients=c, D_zz_coefficients=c)
    _ = pb.D_xx, pb.D_yy, pb.D_zz  # operator pre-compute is not part of the timed path
    return pb, g


def time_cpu(steps, warmup):
    pb, g = cpu_sample_problem()
    for _ in range(warmup):
        cpu_sample_step(pb, g)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_sample_step(pb, g)
        ts.append(time.perf_counter() - t0)
    return 8, ts


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = max((d.get("num_threads", 1) for d in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return os.cpu_count() or 1


SAMPLE_DESC = ("one depth-1 subtree of the workload (8 leaves p=12 q=10: local solves as written with explicit "
               "inverses, 1 oct merge m=100, 1 down pass), NumPy oracle on OpenBLAS")


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_leaves, ts = time_cpu(max(1, args.steps), max(1, min(args.warmup, 1)))
    t = sum(ts) / len(ts)
    val = n_leaves / t
    line = {
        "impl": "reference", "metric": "leaf_solves_per_s_build_plus_solve", "value": val, "unit": "leaves/s",
        "n_gpus": args.gpus, "steps": len(ts), "warmup": max(1, min(args.warmup, 1)), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D variable-coefficient Poisson, uniform octree p={P} q={Q}, DtN, FP64", "L": args.L,
                   "sample": SAMPLE_DESC},
        "cpu_baseline": {"value": val, "unit": "leaves/s", "cores": cpu_threads(), "kind": "port", "sample": SAMPLE_DESC},
        "e2e": {"value": val, "unit": "leaves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm


def run_ours(args, rank, world, local_rank):
    import torch

    import jaxhps_b200 as hps
    from jaxhps_b200 import _lib

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    L = args.L
    dom, c_h, src_h, g_h = synthetic_fields(L)
    n_leaves = dom.n_leaves
    if world > 1:
        from jaxhps_b200 import _dist

        plan = _dist.SubtreePlan(L, rank, world)
        sl = plan.leaf_slice
    else:
        plan, sl = None, slice(0, n_leaves)
    # host (pinned) and resident copies of this rank's inputs
    c_pin = torch.from_numpy(np.ascontiguousarray(c_h[sl])).pin_memory()
    s_pin = torch.from_numpy(np.ascontiguousarray(src_h[sl])).pin_memory()
    g_pin = torch.from_numpy(g_h).pin_memory()
    c_dev, s_dev, g_dev = c_pin.to(dev), s_pin.to(dev), g_pin.to(dev)

    def make_problem(c, s):
        if plan is None:
            return hps.PDEProblem(dom, source=s, D_xx_coefficients=c, D_yy_coefficients=c, D_zz_coefficients=c)
        return _dist.local_problem(dom, plan, source=s, D_xx_coefficients=c, D_yy_coefficients=c, D_zz_coefficients=c)

    pb_res = make_problem(c_dev, s_dev)
    pb_host = make_problem(c_pin, s_pin)

    def step(pb, g, to_host):
        pb.reset()
        if plan is None:
            hps.build_solver(pb, compute_device=dev, host_device=dev)
            u = hps.solve(pb, g, compute_device=dev, host_device=dev)
        else:
            state = _dist.build_solver_sharded(pb, plan, dev)
            u = _dist.solve_sharded(pb, state, plan, g, dev)
        if to_host:
            return u.cpu()
        return u

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(pb, g, to_host, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step(pb, g, to_host)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        # max over ranks
        if dist is not None:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, clocks

    u_chk = None
    for _ in range(args.warmup):
        u_chk = step(pb_res, g_dev, False)
    # self-check against the manufactured solution (not timed)
    u_ref = torch.from_numpy(u_exact(dom.interior_points[sl])).to(dev)
    if u_chk is None:
        u_chk = step(pb_res, g_dev, False)
    err = (u_chk - u_ref).abs().max() / u_ref.abs().max()
    if dist is not None:
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    max_rel_err = float(err)
    del u_chk, u_ref
    # device-resident timing, with the library's kernel timers on (2 event records per GEMM /
    # panel launch; ~0.5% of the step)
    lib.hps_prof_enable(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, wall, clocks = timed(pb_res, g_dev, False, args.steps, sampler)
    pm = (ctypes.c_double * 8)()
    pw = (ctypes.c_double * 8)()
    pl = (ctypes.c_int64 * 8)()
    allk = ctypes.c_int64()
    _lib.check(lib.hps_prof_read(_lib.stream_ptr(), pm, pw, pl, ctypes.byref(allk)), "hps_prof_read")
    lib.hps_prof_enable(0)
    ms_step = ms / args.steps
    value = n_leaves / (ms_step * 1e-3)

    # per-stage breakdown (not part of the timed region): the stage functions called one by one
    stages = None
    if plan is None:
        from jaxhps_b200.down_pass import down_pass_uniform_3D_DtN
        from jaxhps_b200.local_solve import local_solve_stage_uniform_3D_DtN
        from jaxhps_b200.merge import merge_stage_uniform_3D_DtN

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        pb_res.reset()
        torch.cuda.synchronize()
        e0 = ev()
        Y, T, v, h = local_solve_stage_uniform_3D_DtN(pb_res, device=dev, host_device=dev)
        e1 = ev()
        S_lst, gt_lst = merge_stage_uniform_3D_DtN(T, h, L, device=dev, host_device=dev)
        e2 = ev()
        u_s = down_pass_uniform_3D_DtN(g_dev, S_lst, gt_lst, Y, v, device=dev, host_device=dev)
        e3 = ev()
        torch.cuda.synchronize()
        t_loc, t_mrg, t_dwn = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
        down_bytes = 8 * (sum(int(S.numel()) for S in S_lst) + int(Y.numel()))
        leaf_fl = lean_flops(0) * n_leaves  # lean_flops(0) = one leaf
        stages = {
            "local_solve_ms": t_loc, "merge_ms": t_mrg, "down_pass_ms": t_dwn,
            "leaf_solves_per_s_local_solve_stage": n_leaves / (t_loc * 1e-3),
            "local_solve_tflops_lean": leaf_fl / (t_loc * 1e-3) * 1e-12,
            "local_solve_frac_of_fp64_peak": leaf_fl / (t_loc * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS,
            "merge_tflops_lean": (lean_flops(L) - leaf_fl) / (t_mrg * 1e-3) * 1e-12,
            "merge_frac_of_fp64_peak": (lean_flops(L) - leaf_fl) / (t_mrg * 1e-3) * 1e-12 / FP64_PEAK_TFLOPS,
            "down_pass_bytes": down_bytes, "down_pass_GBps": down_bytes / (t_dwn * 1e-3) * 1e-9,
            "down_pass_frac_of_hbm_peak": down_bytes / (t_dwn * 1e-3) * 1e-9 / 6558.1,
        }
        del Y, T, v, h, S_lst, gt_lst, u_s

    # end-to-end: host inputs -> public API -> host result, copies inside the timed region
    step(pb_host, g_pin, True)
    ms_e, _, _ = timed(pb_host, g_pin, True, args.steps)
    e2e_value = n_leaves / (ms_e / args.steps * 1e-3)
    h2d = (c_pin.numel() + s_pin.numel() + g_pin.numel()) * 8
    d2h = (n_leaves // world) * P**3 * 8

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    gemm_ms, gemm_flops, gemm_launches = pm[0], pw[0], pl[0]
    achieved = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    cpu_baseline = None  # timed on rank 0 at N=1 only
    if world == 1:
        cpu_leaves, cpu_ts = time_cpu(1, 1)
        cpu_baseline = {"value": cpu_leaves / cpu_ts[0], "unit": "leaves/s", "cores": cpu_threads(), "kind": "port",
                        "sample": SAMPLE_DESC}
    line = {
        "metric": "leaf_solves_per_s_build_plus_solve", "value": value, "unit": "leaves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D variable-coefficient Poisson, uniform octree p={P} q={Q}, DtN, FP64", "L": L,
                   "n_leaves": n_leaves, "n_bdry": int(g_h.shape[0]),
                   "parallelism": "single GPU" if world == 1 else f"subtree-sharded x{world}; root merge column-sharded by child, root LU distributed by block columns with look-ahead (replicated below n=8192)",
                   "l2_policy": "working set (>=15 GB of operators per step) far exceeds the 126 MB L2; no flush needed"},
        "build_solve_seconds": ms_step * 1e-3,
        "max_rel_error_vs_analytic_solution": max_rel_err,
        "stages": stages,
        "algorithmic_tflop_per_step": lean_flops(L) * 1e-12,
        "step_tflops": lean_flops(L) * 1e-12 / (ms_step * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "leaves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e / args.steps},
        "gpu_launches": int(allk.value),
        "roofline": {"kernel": "hps::gemm_kernel (DMMA m8n8k4 FP64)", "bound": "tensor", "achieved": achieved,
                     "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                     "peak_source": "measured cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r01_fp64_probe.txt); "
                                    "MEASURED_PEAKS.json has no FP64 entry",
                     "launches": int(gemm_launches), "gemm_ms_per_step": gemm_ms / args.steps,
                     "share_of_step": gemm_ms / ms,
                     # dram__bytes_read+write of one 8192^3 launch of this kernel (ncu --set full,
                     # profiles/r01_ncu_gemm_summary.txt); algorithmic operand bytes of that launch: 1.61e9.
                     # The re-reads are tile re-fetches served at < 1 TB/s: the kernel is tensor-bound.
                     "traffic": 32.13e9, "traffic_shape": "8192x8192x8192", "traffic_algorithmic_bytes": 1.61e9,
                     "panel_kernel_ms_per_step": pm[1] / args.steps, "panel_launches": int(pl[1]),
                     "other_kernels_ms_per_step": {name: round(pm[i] / args.steps, 3) for i, name in enumerate(
                         ["gemm", "lu_panel", "trtri", "laswp", "inner_trsm", "merge_gather", "skinny_matvec",
                          "leaf_assemble"])},
                     "hbm_kernels": {"skinny_matvec_GBps": (pw[6] / (pm[6] * 1e-3) * 1e-9) if pm[6] > 0 else None,
                                     "merge_gather_GBps": (pw[5] / (pm[5] * 1e-3) * 1e-9) if pm[5] > 0 else None,
                                     "leaf_assemble_GBps": (pw[7] / (pm[7] * 1e-3) * 1e-9) if pm[7] > 0 else None,
                                     "hbm_peak_GBps": 6558.1}},
        "cpu_baseline": cpu_baseline,
        "wall_s_timed_region": wall,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
