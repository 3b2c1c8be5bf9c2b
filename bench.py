#!/usr/bin/env python
"""bench.py — HPS build_solver + solve on B200, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--L L] [--target-L4 0|1]

One "step" = one full pass of the hot path over the batch of leaves:
local_solve_stage -> merge_stage (all levels) -> down_pass, 3D Poisson-type operator with a
synthetic variable coefficient field, p=12, q=10, FP64 (BASELINE config 3).  At N=1 the tree
has L=3 levels (512 leaves) — the largest configuration whose operators fit one GPU (L=4 needs
a 76 800^2 root merge, 47 GB for D alone plus 94 GB for S).  With 8 ranks the run
additionally times BASELINE's target size L=4 (4096 leaves) and embeds it as ``target_L4``.

value      = leaves per second through build+solve with inputs resident in HBM; timed with the
             library's kernel timers OFF (they are collected in a separate, untimed pass).
e2e        = same through the public API with HOST (pinned) inputs and the solution read back.
e2e_host_resident = e2e with the reference's default ``host_device="cpu"``: every operator the build
             returns (Y, v, S_lst, g_tilde_lst) is copied to the host and back for the solve.
roofline   = DMMA GEMM kernel: algorithmic flops / summed CUDA-event launch time vs the cuBLAS DGEMM
             rate MEASURED IN THIS RUN (torch.matmul fp64 8192^3, burst and 4 s sustained).
cpu_baseline / --impl reference = the NumPy oracle (restatement of the reference's algorithm; the
             reference's JAX runtime is not installable) on ALL host cores, on a bounded sample of the
             workload (one depth-1 subtree: 8 leaves, 1 merge, 1 solve).  ``same_config_sample`` is the
             GPU arm on exactly that 8-leaf sample, so one like-for-like ratio exists.
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
HBM_FALLBACK_GBPS = 6650.0  # B200_PROFILING.md fallback, used only when MEASURED_PEAKS.json is absent
PROF_NAMES = ["gemm", "lu_panel", "trtri", "laswp", "inner_trsm", "merge_gather", "skinny_matvec", "leaf_assemble", "p2p_send",
              "wait_block", "panel_unsort"]


def u_exact(x):
    """Manufactured solution used by every bench problem: u = sin(2 pi x) cos(pi y) exp(z)."""
    return np.sin(2 * np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1]) * np.exp(x[..., 2])


def synthetic_fields(L, p=P):
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


def hbm_peak():
    """(GB/s, source) — the driver-measured copy bandwidth when MEASURED_PEAKS.json travelled with the repo."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return HBM_FALLBACK_GBPS, "fallback (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def committed_gemm_traffic():
    """Per-launch dram__bytes_read+write of the product GEMM kernel from the committed ncu summary
    (profiles/r02_ncu_gemm_summary.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_gemm_summary.json")) as f:
            return json.load(f)
    except Exception:
        return None


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
            return
        # nvidia-smi needs a few hundred ms to initialise NVML, during which kernel launches stall: wait for its first
        # sample here, BEFORE the timed region starts, so that only the steady 200 ms polling runs inside it
        self.first = ""
        try:
            import select

            if select.select([self.proc.stdout], [], [], 3.0)[0]:
                self.first = self.proc.stdout.readline()
        except (OSError, ValueError):
            pass

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
        for line in (getattr(self, "first", "") + out).strip().splitlines():
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
    """The oracle on every host core: torchrun exports OMP_NUM_THREADS=1, so the BLAS pool is widened explicitly."""
    cores = os.cpu_count() or 1
    pb, g = cpu_sample_problem()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        ctx = threadpool_limits(limits=cores)
    except Exception:  # pragma: no cover
        import contextlib

        ctx, threadpool_info = contextlib.nullcontext(), None
    with ctx:
        for _ in range(warmup):
            cpu_sample_step(pb, g)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            cpu_sample_step(pb, g)
            ts.append(time.perf_counter() - t0)
        used = cores
        if threadpool_info is not None:
            used = max((d.get("num_threads", 1) for d in threadpool_info()), default=cores)
    return 8, ts, int(used)


SAMPLE_DESC = ("one depth-1 subtree of the workload (8 leaves p=12 q=10: local solves as written with explicit "
               "inverses, 1 oct merge m=100, 1 down pass), NumPy oracle on OpenBLAS, all host cores")
WORKLOAD = f"3D variable-coefficient Poisson, uniform octree p={P} q={Q}, DtN, FP64"


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_leaves, ts, cores = time_cpu(max(1, args.steps), max(0, args.warmup))
    t = sum(ts) / len(ts)
    val = n_leaves / t
    line = {
        "impl": "reference", "metric": "leaf_solves_per_s_build_plus_solve", "value": val, "unit": "leaves/s",
        "n_gpus": args.gpus, "steps": len(ts), "warmup": max(0, args.warmup), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "L": args.L, "sample": SAMPLE_DESC,
                   "note": "the CPU arm times the 8-leaf sample, not the full tree (L=3 is > 1 h of CPU); the GPU arm's "
                           "`same_config_sample` times the identical sample for a like-for-like ratio"},
        "cpu_baseline": {"value": val, "unit": "leaves/s", "cores": cores, "kind": "port", "sample": SAMPLE_DESC},
        "e2e": {"value": val, "unit": "leaves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm


def measure_fp64_peak(torch, dev, sustain_s=4.0):
    """cuBLAS DGEMM 8192^3 through torch.matmul: best of 10 (burst) and back to back for `sustain_s` (sustained).
    Only the roofline denominator — nothing on the product path calls cuBLAS."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    fl = 2.0 * n**3
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    one = best * 1e-3
    reps = max(10, int(sustain_s / one))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    sustained = fl * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    del a, b, c
    return fl / one * 1e-12, sustained


class Runner:
    """One problem size on this rank: resident and pinned-host copies of the inputs, the step function."""

    def __init__(self, torch, hps, L, rank, world, dev, dist):
        self.torch, self.hps, self.L, self.rank, self.world, self.dev, self.dist = torch, hps, L, rank, world, dev, dist
        dom, c_h, src_h, g_h = synthetic_fields(L)
        self.dom, self.n_leaves, self.n_bdry = dom, dom.n_leaves, int(g_h.shape[0])
        if world > 1:
            from jaxhps_b200 import _dist

            self._dist = _dist
            self.plan = _dist.SubtreePlan(L, rank, world)
            sl = self.plan.leaf_slice
        else:
            self._dist, self.plan, sl = None, None, slice(0, dom.n_leaves)
        self.sl = sl
        self.c_pin = torch.from_numpy(np.ascontiguousarray(c_h[sl])).pin_memory()
        self.s_pin = torch.from_numpy(np.ascontiguousarray(src_h[sl])).pin_memory()
        self.g_pin = torch.from_numpy(g_h).pin_memory()
        self.c_dev, self.s_dev, self.g_dev = self.c_pin.to(dev), self.s_pin.to(dev), self.g_pin.to(dev)
        self.pb_res = self.make_problem(self.c_dev, self.s_dev)
        self.pb_host = self.make_problem(self.c_pin, self.s_pin)

    def make_problem(self, c, s):
        if self.plan is None:
            return self.hps.PDEProblem(self.dom, source=s, D_xx_coefficients=c, D_yy_coefficients=c, D_zz_coefficients=c)
        return self._dist.local_problem(self.dom, self.plan, source=s, D_xx_coefficients=c, D_yy_coefficients=c,
                                        D_zz_coefficients=c)

    def step_factored(self, g):
        """One build+solve in the opt-in factored-root (S-free) mode of the sharded driver (any world size)."""
        from jaxhps_b200 import _dist

        if self.plan is not None:
            plan, pb = self.plan, self.pb_res
        else:
            if not hasattr(self, "_plan1"):
                self._plan1 = _dist.SubtreePlan(self.L, 0, 1)
                self._pb1 = _dist.local_problem(self.dom, self._plan1, source=self.s_dev, D_xx_coefficients=self.c_dev,
                                                D_yy_coefficients=self.c_dev, D_zz_coefficients=self.c_dev)
            plan, pb = self._plan1, self._pb1
        pb.reset()
        state = _dist.build_solver_sharded(pb, plan, self.dev, root_mode="factored")
        return _dist.solve_sharded(pb, state, plan, g, self.dev)

    def timed_factored(self, steps):
        torch = self.torch
        u = self.step_factored(self.g_dev)  # warm-up
        err = self.error_vs_analytic(u)
        del u
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step_factored(self.g_dev)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t[0])
        return {"ms_per_step": ms / steps, "leaves_per_s": self.n_leaves / (ms / steps * 1e-3), "steps": steps, "warmup": 1,
                "max_rel_error_vs_analytic_solution": err,
                "note": "opt-in root_mode='factored': the root S = -D^-1 C is not formed (the reference's S_lst[-1] does not "
                        "exist in this mode); every solve applies D^-1 from the kept LU factors instead"}

    def step(self, pb, g, to_host, host_device=None):
        dev = self.dev
        pb.reset()
        if self.plan is None:
            hd = dev if host_device is None else host_device
            self.hps.build_solver(pb, compute_device=dev, host_device=hd)
            u = self.hps.solve(pb, g, compute_device=dev, host_device=hd)
        else:
            state = self._dist.build_solver_sharded(pb, self.plan, dev)
            u = self._dist.solve_sharded(pb, state, self.plan, g, dev)
        if to_host and hasattr(u, "cpu"):
            return u.cpu()
        return u

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, pb, g, to_host, steps, sampler=None, host_device=None):
        torch = self.torch
        if sampler:
            sampler.start()  # returns once the sampler is polling; the other ranks wait in the barrier
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            self.step(pb, g, to_host, host_device)
        e1.record()
        self.barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if self.dist is not None:  # max over ranks
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, clocks

    def error_vs_analytic(self, u):
        torch = self.torch
        u_ref = torch.from_numpy(u_exact(self.dom.interior_points[self.sl])).to(self.dev)
        err = (u - u_ref).abs().max() / u_ref.abs().max()
        if self.dist is not None:
            self.dist.all_reduce(err, op=self.dist.ReduceOp.MAX)
        return float(err)


def read_prof(lib, _lib):
    pm = (ctypes.c_double * 16)()
    pw = (ctypes.c_double * 16)()
    pl = (ctypes.c_int64 * 16)()
    allk = ctypes.c_int64()
    _lib.check(lib.hps_prof_read(_lib.stream_ptr(), pm, pw, pl, ctypes.byref(allk)), "hps_prof_read")
    return list(pm), list(pw), list(pl), int(allk.value)


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

    # ---- roofline denominators measured in this run (rank 0's GPU; the other ranks wait at the next barrier) ----
    fp64_burst, fp64_sustained = measure_fp64_peak(torch, dev) if rank == 0 else (None, None)
    hbm_gbps, hbm_src = hbm_peak()

    L = args.L
    R = Runner(torch, hps, L, rank, world, dev, dist)
    n_leaves, n_bdry = R.n_leaves, R.n_bdry

    u_chk = None
    for _ in range(args.warmup):
        u_chk = R.step(R.pb_res, R.g_dev, False)
    if u_chk is None:
        u_chk = R.step(R.pb_res, R.g_dev, False)
    max_rel_err = R.error_vs_analytic(u_chk)  # self-check against the manufactured solution (not timed)
    del u_chk

    # ---- headline: device-resident inputs, library kernel timers OFF ----
    lib.hps_prof_enable(0)  # also resets the launch counter
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, wall, clocks = R.timed(R.pb_res, R.g_dev, False, args.steps, sampler)
    _, _, _, launches_timed = read_prof(lib, _lib)
    ms_step = ms / args.steps
    value = n_leaves / (ms_step * 1e-3)

    # ---- separate, untimed pass with the per-category kernel timers on (2 event records per launch) ----
    lib.hps_prof_enable(1)
    R.barrier()
    R.step(R.pb_res, R.g_dev, False)
    R.barrier()
    pm, pw, pl, _ = read_prof(lib, _lib)
    lib.hps_prof_enable(0)

    # ---- sharded runs: stage boundaries of one more untimed build on rank 0 (CUDA events; developer breakdown) ----
    sharded_stages = None
    if R.plan is not None:
        R._dist.STAGE_TIMING = True
        R.pb_res.reset()
        st_dbg = R._dist.build_solver_sharded(R.pb_res, R.plan, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        R._dist.solve_sharded(R.pb_res, st_dbg, R.plan, R.g_dev, dev)
        e1.record()
        torch.cuda.synchronize()
        sharded_stages = dict(st_dbg.timing or {})
        sharded_stages["solve"] = e0.elapsed_time(e1)
        R._dist.STAGE_TIMING = False
        del st_dbg
        R.barrier()

    # ---- per-stage breakdown (single GPU; not part of the timed region): the stage functions one by one ----
    stages = None
    fp64_peak_for_stages = fp64_sustained
    if R.plan is None:
        from jaxhps_b200.down_pass import down_pass_uniform_3D_DtN
        from jaxhps_b200.local_solve import local_solve_stage_uniform_3D_DtN
        from jaxhps_b200.merge import merge_stage_uniform_3D_DtN

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        passes = []
        for _ in range(2):  # the first pass may pay allocator growth after the profiled step; the faster one is reported
            R.pb_res.reset()
            Y = T = v = h = S_lst = gt_lst = u_s = None
            torch.cuda.synchronize()
            e0 = ev()
            Y, T, v, h = local_solve_stage_uniform_3D_DtN(R.pb_res, device=dev, host_device=dev)
            e1 = ev()
            S_lst, gt_lst = merge_stage_uniform_3D_DtN(T, h, L, device=dev, host_device=dev)
            e2 = ev()
            u_s = down_pass_uniform_3D_DtN(R.g_dev, S_lst, gt_lst, Y, v, device=dev, host_device=dev)
            e3 = ev()
            torch.cuda.synchronize()
            passes.append((e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)))
        t_loc, t_mrg, t_dwn = (min(p[i] for p in passes) for i in range(3))
        down_bytes = 8 * (sum(int(S.numel()) for S in S_lst) + int(Y.numel()))
        leaf_fl = lean_flops(0) * n_leaves  # lean_flops(0) = one leaf
        stages = {
            "local_solve_ms": t_loc, "merge_ms": t_mrg, "down_pass_ms": t_dwn, "stage_passes_ms": passes,
            "leaf_solves_per_s_local_solve_stage": n_leaves / (t_loc * 1e-3),
            "local_solve_tflops_lean": leaf_fl / (t_loc * 1e-3) * 1e-12,
            "local_solve_frac_of_fp64_peak": leaf_fl / (t_loc * 1e-3) * 1e-12 / fp64_peak_for_stages,
            "merge_tflops_lean": (lean_flops(L) - leaf_fl) / (t_mrg * 1e-3) * 1e-12,
            "merge_frac_of_fp64_peak": (lean_flops(L) - leaf_fl) / (t_mrg * 1e-3) * 1e-12 / fp64_peak_for_stages,
            "down_pass_bytes": down_bytes, "down_pass_GBps": down_bytes / (t_dwn * 1e-3) * 1e-9,
            "down_pass_frac_of_hbm_peak": down_bytes / (t_dwn * 1e-3) * 1e-9 / hbm_gbps,
        }
        del Y, T, v, h, S_lst, gt_lst, u_s

    # ---- end-to-end: host inputs -> public API -> host result, copies inside the timed region ----
    R.step(R.pb_host, R.g_pin, True)
    sampler_e = ClockSampler(local_rank) if rank == 0 else None  # same conditions as the headline (nvidia-smi polling costs ~1 %)
    ms_e, _, clocks_e = R.timed(R.pb_host, R.g_pin, True, args.steps, sampler_e)
    e2e_value = n_leaves / (ms_e / args.steps * 1e-3)
    h2d = (R.c_pin.numel() * 3 + R.s_pin.numel() + R.g_pin.numel()) * 8  # c is passed (and copied) as D_xx, D_yy and D_zz
    d2h = (n_leaves // world) * P**3 * 8

    # ---- end-to-end with the reference's default host_device="cpu" (single GPU): operators round-trip the host ----
    e2e_host = None
    if R.plan is None and args.host_resident:
        torch.cuda.empty_cache()
        R.step(R.pb_host, R.g_pin, True, host_device="cpu")  # warm-up: the pinned result buffers are allocated once
        ms_h, _, _ = R.timed(R.pb_host, R.g_pin, True, 1, host_device="cpu")
        op_bytes = 8 * (n_leaves * P**3 * 6 * Q * Q + n_leaves * P**3)
        m = Q * Q
        for level in range(L, 0, -1):
            op_bytes += 8 * 8 ** (level - 1) * (12 * m) * (24 * m + 1)
            m *= 4
        e2e_host = {"value": n_leaves / (ms_h * 1e-3), "unit": "leaves/s", "ms_per_step": ms_h, "steps": 1,
                    "d2h_operator_bytes_per_step": int(op_bytes), "h2d_operator_bytes_per_step": int(op_bytes),
                    "note": "build_solver(host_device='cpu') returns Y, v, S_lst, g_tilde_lst as NumPy arrays like the "
                            "reference's default (in page-locked host memory); solve() copies them back; 1 warm-up step"}

    # ---- the opt-in factored-root (S-free) mode, same problem (not the headline: S_lst[-1] is not produced) ----
    factored = None
    if args.factored:
        try:
            factored = R.timed_factored(args.steps)
        except Exception as e:  # the mode needs the P2P segment (CUDA IPC); report instead of failing the bench
            factored = {"unavailable": repr(e)[:300]}

    # ---- like-for-like with the CPU arm: the identical 8-leaf depth-1 sample on the GPU (single GPU) ----
    same_sample = None
    if world == 1:
        R1 = Runner(torch, hps, 1, 0, 1, dev, None)
        for _ in range(3):
            R1.step(R1.pb_host, R1.g_pin, True)
        ms_1, _, _ = R1.timed(R1.pb_host, R1.g_pin, True, 10)
        ms_1r, _, _ = R1.timed(R1.pb_res, R1.g_dev, False, 10)
        same_sample = {"sample": "8 leaves p=12 q=10, depth 1 (the CPU arm's sample)", "e2e_leaves_per_s": 8 / (ms_1 / 10 * 1e-3),
                       "resident_leaves_per_s": 8 / (ms_1r / 10 * 1e-3), "e2e_ms_per_step": ms_1 / 10}
        del R1

    # ---- the other BASELINE configs (1, 2, 4, 5) through the public API, single GPU, after everything timed above ----
    other_configs = None
    if world == 1 and args.other_configs:
        try:
            del R
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import other_configs as oc

            other_configs = oc.run_all(hps)
        except Exception as e:  # noqa: BLE001
            other_configs = {"unavailable": repr(e)[:300]}

    # ---- BASELINE's target size on 8 GPUs: L=4, 4096 leaves ----
    target_L4 = None
    if world == 8 and args.target_L4 and L != 4:
        try:  # a failure at the target size must not cost the L=3 line of this run
            del R
            torch.cuda.empty_cache()
            R4 = Runner(torch, hps, 4, rank, world, dev, dist)
            u4 = R4.step(R4.pb_res, R4.g_dev, False)  # warm-up
            err4 = R4.error_vs_analytic(u4)
            del u4
            lib.hps_prof_enable(0)
            ms4, wall4, _ = R4.timed(R4.pb_res, R4.g_dev, False, 2)
            lib.hps_prof_enable(1)
            R4.barrier()
            R4.step(R4.pb_res, R4.g_dev, False)
            R4.barrier()
            pm4, pw4, pl4, _ = read_prof(lib, _lib)
            lib.hps_prof_enable(0)
            target_L4 = {"L": 4, "n_leaves": R4.n_leaves, "n_gpus": world, "steps": 2, "warmup": 1,
                         "ms_per_step": ms4 / 2, "build_solve_seconds": ms4 / 2 * 1e-3,
                         "leaves_per_s": R4.n_leaves / (ms4 / 2 * 1e-3),
                         "max_rel_error_vs_analytic_solution": err4,
                         "algorithmic_tflop_per_step": lean_flops(4) * 1e-12,
                         "step_tflops_per_gpu": lean_flops(4) * 1e-12 / (ms4 / 2 * 1e-3) / world,
                         "gemm_tflops_rank0": (pw4[0] / (pm4[0] * 1e-3) * 1e-12) if pm4[0] > 0 else None,
                         "kernel_ms_rank0": {name: round(pm4[i], 2) for i, name in enumerate(PROF_NAMES)}}
            if args.factored:
                try:
                    target_L4["factored_root"] = R4.timed_factored(2)
                except Exception as e:
                    target_L4["factored_root"] = {"unavailable": repr(e)[:300]}
            del R4

        except Exception as e:
            target_L4 = {"L": 4, "unavailable": repr(e)[:300]}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    gemm_ms, gemm_flops, gemm_launches = pm[0], pw[0], pl[0]
    achieved = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    cpu_baseline = None  # timed on rank 0 at N=1 only
    if world == 1:
        cpu_leaves, cpu_ts, cores = time_cpu(1, 1)
        cpu_baseline = {"value": cpu_leaves / cpu_ts[0], "unit": "leaves/s", "cores": cores, "kind": "port",
                        "sample": SAMPLE_DESC}
    traffic = committed_gemm_traffic()
    line = {
        "metric": "leaf_solves_per_s_build_plus_solve", "value": value, "unit": "leaves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "L": L, "n_leaves": n_leaves, "n_bdry": n_bdry,
                   "parallelism": "single GPU" if world == 1 else f"subtree-sharded x{world}; root merge column-sharded by child, root LU distributed by block columns with look-ahead (replicated below n=8192)",
                   "l2_policy": "working set (>=15 GB of operators per step) far exceeds the 126 MB L2; no flush needed"},
        "build_solve_seconds": ms_step * 1e-3,
        "max_rel_error_vs_analytic_solution": max_rel_err,
        "stages": stages,
        "sharded_stages_ms_rank0": sharded_stages,
        "algorithmic_tflop_per_step": lean_flops(L) * 1e-12,
        "executed_gemm_tflop_per_step": gemm_flops * 1e-12 * (world if world > 1 else 1),
        "executed_note": "sum of 2MNK over the DMMA GEMM launches of one step (rank 0's share x ranks when sharded); below the "
                         "algorithmic count because the forward substitutions skip the structurally-zero rows of -C "
                         "(DESIGN 4c) - stage fractions computed from the algorithmic count can therefore exceed 1",
        "step_tflops": lean_flops(L) * 1e-12 / (ms_step * 1e-3),
        "step_frac_of_fp64_peak": lean_flops(L) * 1e-12 / (ms_step * 1e-3) / world / fp64_sustained,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "leaves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e / args.steps, "clocks": clocks_e},
        "e2e_host_resident": e2e_host,
        "factored_root": factored,
        "same_config_sample": same_sample,
        "target_L4": target_L4,
        "other_configs": other_configs,
        "gpu_launches": launches_timed,
        "roofline": {"kernel": "hps::gemmk::gemm_kernel_hoist (DMMA m8n8k4 FP64)", "bound": "tensor", "achieved": achieved,
                     "peak": fp64_sustained, "unit": "TFLOP/s", "frac": achieved / fp64_sustained,
                     "peak_source": "measured in this run: cuBLAS DGEMM 8192^3 via torch.matmul, back to back for 4 s "
                                    "(sustained; the kernel is timed inside a long step)",
                     "peak_burst": fp64_burst, "frac_of_burst": achieved / fp64_burst,
                     "launches": int(gemm_launches), "gemm_ms_per_step": gemm_ms,
                     "share_of_step": gemm_ms / ms_step,
                     "timers": "separate untimed pass after the headline (kernel timers are off in the timed region)",
                     "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                     "traffic_detail": traffic,
                     "other_kernels_ms_per_step": {name: round(pm[i], 3) for i, name in enumerate(PROF_NAMES)},
                     "lu_block_column_launches": int(pl[1]),
                     "hbm_kernels": {"skinny_matvec_GBps": (pw[6] / (pm[6] * 1e-3) * 1e-9) if pm[6] > 0 else None,
                                     "merge_gather_GBps": (pw[5] / (pm[5] * 1e-3) * 1e-9) if pm[5] > 0 else None,
                                     "leaf_assemble_GBps": (pw[7] / (pm[7] * 1e-3) * 1e-9) if pm[7] > 0 else None,
                                     "panel_unsort_GBps": (pw[10] / (pm[10] * 1e-3) * 1e-9) if pm[10] > 0 else None,
                                     "hbm_peak_GBps": hbm_gbps, "hbm_peak_source": hbm_src}},
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
    ap.add_argument("--target-L4", dest="target_L4", type=int, default=1,
                    help="with 8 ranks: also time BASELINE's target size L=4 (1 warm-up + 2 steps) and embed it")
    ap.add_argument("--factored", type=int, default=1,
                    help="also time the opt-in factored-root (S-free) mode of the sharded driver")
    ap.add_argument("--other-configs", dest="other_configs", type=int, default=1,
                    help="single GPU: also run BASELINE configs 1, 2, 4 and 5 once (about 15 s) and embed their numbers")
    ap.add_argument("--host-resident", dest="host_resident", type=int, default=1,
                    help="single GPU: also time one step with the reference-default host_device='cpu'")
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
