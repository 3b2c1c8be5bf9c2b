"""Time hps_lu_solve (factorisation only, or with right-hand sides) through the C ABI and split the time by
kernel category (developer tool).  usage: bench_lu.py n batch [n_rhs_cols] [reps]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from jaxhps_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
n, batch = int(sys.argv[1]), int(sys.argv[2])
ncols = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
A0 = torch.randn(batch, n, n, dtype=torch.float64, device=dev) + 0.0
if os.environ.get("HPS_LU_FORCE_SPEC") == "1":  # a matrix that partial pivoting leaves alone, like the merges' D
    A0 += 4.0 * n**0.5 * torch.eye(n, dtype=torch.float64, device=dev)
B0 = torch.randn(batch, n, max(ncols, 1), dtype=torch.float64, device=dev)
need = ctypes.c_size_t()
_lib.check(lib.hps_lu_solve_workspace(batch, n, ctypes.byref(need)), "ws")
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
info = torch.zeros(batch, dtype=torch.int32, device=dev)
names = ["gemm", "lu_panel", "trtri", "laswp", "inner_trsm", "merge_gather", "skinny", "assemble", "p2p_send"]


def run(A, B):
    k = 1 if ncols else 0
    ptrs = (ctypes.c_void_p * 1)(B.data_ptr())
    lds = (ctypes.c_int64 * 1)(B.shape[2])
    strides = (ctypes.c_int64 * 1)(B.shape[1] * B.shape[2])
    ncs = (ctypes.c_int * 1)(B.shape[2])
    _lib.check(lib.hps_lu_solve(_lib.stream_ptr(), batch, n, A.data_ptr(), n, n * n, k, ptrs, lds, strides, ncs,
                                ws.data_ptr(), ws.numel(), info.data_ptr()), "lu_solve")


for it in range(reps + 1):
    A, B = A0.clone(), B0.clone()
    torch.cuda.synchronize()
    if it == reps:
        lib.hps_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(A, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = batch * (2 / 3 * n**3 + 2.0 * n * n * ncols)
    print(f"n={n} batch={batch} ncols={ncols} iter {it}: {ms:.3f} ms  {fl / ms * 1e-9:.2f} TF/s" + ("  (profiler on)" if it == reps else ""))
pm = (ctypes.c_double * 16)()
pw = (ctypes.c_double * 16)()
pl = (ctypes.c_int64 * 16)()
allk = ctypes.c_int64()
lib.hps_prof_read(_lib.stream_ptr(), pm, pw, pl, ctypes.byref(allk))
lib.hps_prof_enable(0)
print("   per category ms:", {nm: round(pm[i], 3) for i, nm in enumerate(names) if pl[i]}, "launches", {nm: pl[i] for i, nm in enumerate(names) if pl[i]})
if ncols:
    X = B
    R = torch.bmm(A0, X) - B0
    print("   residual", float(R.abs().max() / (A0.abs().max() * X.abs().max() * n)))
else:
    assert int(info.abs().max()) == 0, info
