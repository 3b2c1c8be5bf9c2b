"""Where one L=3 build+solve spends its time (developer tool, 1 GPU): per-launch timeline of the library profiler,
aggregated (a) per stream and category, (b) per GEMM shape (M, N, K, batch) with the achieved TFLOP/s, (c) idle gaps of
the caller's stream.  usage: python tools/profile_step.py [L]  ->  gpurun_out/step_profile_L{L}.txt"""
import ctypes
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import jaxhps_b200 as hps  # noqa: E402
from jaxhps_b200 import _lib  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
lib = _lib.load()
run = bench.Runner(torch, hps, L, 0, 1, dev, None)
for _ in range(2):
    run.step(run.pb_res, run.g_dev, False)
torch.cuda.synchronize()
lib.hps_prof_enable(1)
run.step(run.pb_res, run.g_dev, False)
torch.cuda.synchronize()
cap = 400000
t0, t1, wk = (ctypes.c_double * cap)(), (ctypes.c_double * cap)(), (ctypes.c_double * cap)()
cat, sid, dims = (ctypes.c_int * cap)(), (ctypes.c_int * cap)(), (ctypes.c_int * (4 * cap))()
n = ctypes.c_int64()
lib.hps_prof_timeline(t0, t1, wk, cat, sid, dims, cap, ctypes.byref(n))
lib.hps_prof_enable(0)
names = ["gemm", "lu_panel", "trtri", "laswp", "inner_trsm", "merge_gather", "skinny", "assemble", "p2p_send", "wait_block", "panel_unsort"]
N = n.value
out = []
span = max(t1[i] for i in range(N)) - min(t0[i] for i in range(N))
out.append(f"L={L}: {N} profiled launches, span {span:.1f} ms (profiler on: two event records per launch)")
busy = defaultdict(lambda: [0, 0.0])
for i in range(N):
    b = busy[(sid[i], names[cat[i]])]
    b[0] += 1
    b[1] += t1[i] - t0[i]
out.append("per (stream, category): launches, ms")
for k, v in sorted(busy.items()):
    out.append(f"   stream {k[0]} {k[1]:12s} {v[0]:6d} {v[1]:9.2f}")
shapes = defaultdict(lambda: [0, 0.0, 0.0])
for i in range(N):
    if cat[i] == 0:
        s = shapes[(dims[4 * i], dims[4 * i + 1], dims[4 * i + 2], dims[4 * i + 3])]
        s[0] += 1
        s[1] += t1[i] - t0[i]
        s[2] += wk[i]
tot_ms = sum(v[1] for v in shapes.values())
tot_fl = sum(v[2] for v in shapes.values())
out.append(f"GEMM: {tot_ms:.1f} ms, {tot_fl * 1e-12:.2f} TFLOP, {tot_fl / tot_ms * 1e-9:.2f} TF/s; by K class and by shape (top 40 by time):")
kcls = defaultdict(lambda: [0, 0.0, 0.0])
for (M, Nn, K, B), v in shapes.items():
    key = "K<=32" if K <= 32 else ("K<=128" if K <= 128 else ("K<=512" if K <= 512 else "K>512"))
    kcls[key][0] += v[0]
    kcls[key][1] += v[1]
    kcls[key][2] += v[2]
for k, v in sorted(kcls.items()):
    out.append(f"   {k:8s} launches {v[0]:5d}  {v[1]:8.2f} ms  {v[2] * 1e-12:7.3f} TFLOP  {v[2] / max(v[1], 1e-9) * 1e-9:6.2f} TF/s")
for (M, Nn, K, B), v in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:40]:
    out.append(f"   M={M:6d} N={Nn:6d} K={K:6d} batch={B:4d}  x{v[0]:4d}  {v[1]:8.2f} ms  {v[2] / max(v[1], 1e-9) * 1e-9:6.2f} TF/s")
# idle gaps on the stream that carries most GEMM time
main = max({s for s, _ in busy}, key=lambda s: busy[(s, "gemm")][1] if (s, "gemm") in busy else 0.0)
ev = sorted((t0[i], t1[i], names[cat[i]]) for i in range(N) if sid[i] == main)
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    if b[0] - a[1] > 0.02:
        gaps.append((b[0] - a[1], a[1], a[2], b[2]))
out.append(f"stream {main}: busy {sum(e[1] - e[0] for e in ev):.1f} ms of {ev[-1][1] - ev[0][0]:.1f}; gaps > 20 us: {len(gaps)}, total {sum(g[0] for g in gaps):.1f} ms; largest:")
for g in sorted(gaps, reverse=True)[:25]:
    out.append(f"   {g[0]:7.3f} ms at {g[1]:9.2f}  after {g[2]} before {g[3]}")
# stage windows: the leaf stage ends at the first merge_gather; every later merge_gather starts a merge level
cuts = sorted(t0[i] for i in range(N) if names[cat[i]] == "merge_gather")
edges = [min(t0[i] for i in range(N))] + cuts + [max(t1[i] for i in range(N)) + 1e-3]
for w in range(len(edges) - 1):
    a, b = edges[w], edges[w + 1]
    seg = defaultdict(lambda: [0, 0.0, 0.0])
    for i in range(N):
        if a <= t0[i] < b:
            v = seg[(sid[i], names[cat[i]])]
            v[0] += 1
            v[1] += t1[i] - t0[i]
            v[2] += wk[i]
    out.append(f"window {w} [{a:.1f}, {b:.1f}) ms = {b - a:.1f} ms: " + ", ".join(
        f"s{k[0]}/{k[1]} x{v[0]} {v[1]:.1f}ms" + (f" {v[2] / max(v[1], 1e-9) * 1e-9:.1f}TF/s" if k[1] == "gemm" else "")
        for k, v in sorted(seg.items())))
    if w == 0:
        for i in sorted((i for i in range(N) if a <= t0[i] < b and t1[i] - t0[i] > 0.5), key=lambda i: t0[i]):
            out.append(f"      {t0[i]:8.2f} +{t1[i] - t0[i]:6.2f} ms s{sid[i]} {names[cat[i]]:10s} dims {dims[4*i]},{dims[4*i+1]},{dims[4*i+2]},{dims[4*i+3]}")
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/step_profile_L{L}.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
