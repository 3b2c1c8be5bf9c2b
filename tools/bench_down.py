"""Down-pass / leaf-stage micro-benchmarks (developer tool)."""
import sys, ctypes
import numpy as np, torch
sys.path.insert(0, ".")
import jaxhps_b200 as hps
from jaxhps_b200 import _lib
from jaxhps_b200.local_solve import local_solve_stage_uniform_3D_DtN
lib = _lib.load(); dev = torch.device("cuda:0")
def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
# matvec shapes of the L=3 down pass
for (M, K, batch) in ((19200, 38400, 1), (4800, 9600, 8), (1200, 2400, 64), (1728, 600, 512)):
    A = torch.randn(batch, M, K, dtype=torch.float64, device=dev); x = torch.randn(batch, K, 1, dtype=torch.float64, device=dev)
    c = torch.randn(batch, M, 1, dtype=torch.float64, device=dev); out = torch.empty_like(c)
    f = lambda: lib.hps_leaf_apply(_lib.stream_ptr(), batch, M, K, 1, A.data_ptr(), x.data_ptr(), c.data_ptr(), out.data_ptr())
    ms = ev(f); ref = torch.bmm(A, x) + c
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"matvec M={M} K={K} b={batch}: {ms:.3f} ms  {8*M*K*batch/ms*1e-6:.0f} GB/s  err={err:.1e}", flush=True)
    del A
# leaf stage at p=12, 512 leaves
rng = np.random.default_rng(0)
dom = hps.Domain(12, 10, hps.DiscretizationNode3D(0., 1., 0., 1., 0., 1.), 3)
shp = dom.interior_points[..., 0].shape
c = torch.from_numpy(1 + 0.1 * rng.normal(size=shp)).to(dev); src = torch.from_numpy(rng.normal(size=shp)).to(dev)
pb = hps.PDEProblem(dom, source=src, D_xx_coefficients=c, D_yy_coefficients=c, D_zz_coefficients=c)
lib.hps_prof_enable(1)
ms = ev(lambda: local_solve_stage_uniform_3D_DtN(pb, device=dev, host_device=dev), 3)
pm = (ctypes.c_double * 16)(); pw = (ctypes.c_double * 16)(); pl = (ctypes.c_int64 * 16)(); al = ctypes.c_int64()
lib.hps_prof_read(_lib.stream_ptr(), pm, pw, pl, ctypes.byref(al))
print(f"leaf stage 512 leaves: {ms:.1f} ms  ({512/ms*1e3:.0f} leaves/s, {512*3.99e9/ms*1e-9:.1f} TF/s lean)")
print("per-category ms over 4 calls:", {n: round(pm[i], 1) for i, n in enumerate(["gemm", "panel", "trtri", "laswp", "inner", "gather", "skinny", "assemble"])})
