"""Distributed LU micro-benchmark (developer tool; run under torchrun): hps_lu_dist_run on a random n x n matrix with
`ncols` right-hand-side columns per rank, per-category kernel times of rank 0.
usage: torchrun ... tools/bench_dist_lu.py n ncols [reps]"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from jaxhps_b200 import _dist, _lib

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
n, ncols = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = torch.Generator(device="cpu").manual_seed(0)
A0 = torch.randn(n, n, dtype=torch.float64, generator=g).to(dev)  # same matrix on every rank
A0 += 4.0 * n**0.5 * torch.eye(n, dtype=torch.float64, device=dev)  # like the merges' D: partial pivoting leaves it alone
B0 = torch.randn(n, ncols, dtype=torch.float64, device=dev)
comm = _dist.P2PComm.get(_lib, dev, rank, world, None)
comm.ensure(comm.lu_segment_bytes(n))
names = ["gemm", "lu_panel", "trtri", "laswp", "inner_trsm", "merge_gather", "skinny", "assemble", "p2p_send", "wait_block", "panel_unsort"]
for it in range(reps + 1):
    B = B0.clone()
    _dist._copy_into_segment(comm.matrix_ptr(n), A0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if it == reps:
        lib.hps_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _dist.p2p_lu_solve(_lib, dev, comm, n, [B], None)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} n={n} ncols/rank={ncols} send={os.environ.get('HPS_DIST_SEND', 'copy')} iter {it}: {float(t):.2f} ms" + ("  (profiler on)" if it == reps else ""), flush=True)
pm, pw, pl = (ctypes.c_double * 16)(), (ctypes.c_double * 16)(), (ctypes.c_int64 * 16)()
allk = ctypes.c_int64()
lib.hps_prof_read(_lib.stream_ptr(), pm, pw, pl, ctypes.byref(allk))
res = float((A0 @ B - B0).abs().max() / (A0.abs().max() * B.abs().max() * n))
if rank == 0:
    print("   rank 0 per category ms:", {nm: round(pm[i], 2) for i, nm in enumerate(names) if pl[i]},
          "launches", {nm: pl[i] for i, nm in enumerate(names) if pl[i]}, "p2p GB/s", round(pw[8] / max(pm[8], 1e-9) * 1e-6, 1),
          "residual", res, flush=True)
# per-launch timeline of the profiled iteration, every rank: gpurun_out/dist_lu_timeline_w{world}_r{rank}.csv
cap = 200000
t0, t1, wk = (ctypes.c_double * cap)(), (ctypes.c_double * cap)(), (ctypes.c_double * cap)()
cat, sid, dims = (ctypes.c_int * cap)(), (ctypes.c_int * cap)(), (ctypes.c_int * (4 * cap))()
nrec = ctypes.c_int64()
lib.hps_prof_timeline(t0, t1, wk, cat, sid, dims, cap, ctypes.byref(nrec))
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/dist_lu_timeline_w{world}_r{rank}.csv", "w") as f:
    f.write("t0_ms,t1_ms,category,stream\n")
    for i in range(nrec.value):
        f.write(f"{t0[i]:.4f},{t1[i]:.4f},{names[cat[i]]},{sid[i]}\n")
if rank == 0:
    busy = {}
    for i in range(nrec.value):
        busy.setdefault((sid[i], names[cat[i]]), [0, 0.0])
        busy[(sid[i], names[cat[i]])][0] += 1
        busy[(sid[i], names[cat[i]])][1] += t1[i] - t0[i]
    print("   rank 0 (stream, category): launches, ms:", {k: (v[0], round(v[1], 2)) for k, v in sorted(busy.items())}, flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
