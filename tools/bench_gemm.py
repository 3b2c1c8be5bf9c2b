"""GEMM / LU micro-benchmarks for the shapes that dominate the p=12 path (developer tool)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from jaxhps_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
def ev(fn, n=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
def gemm(M, N, K, batch=1, beta=1.0, lda=None, ldb=None, ldc=None, sA=None):
    lda = lda or K; ldb = ldb or N; ldc = ldc or N
    A = torch.randn(batch if sA != 0 else 1, M, lda, dtype=torch.float64, device=dev)
    B = torch.randn(batch, K, ldb, dtype=torch.float64, device=dev)
    C = torch.randn(batch, M, ldc, dtype=torch.float64, device=dev)
    f = lambda: lib.hps_dgemm_strided_batched(_lib.stream_ptr(), M, N, K, -1.0, A.data_ptr(), lda, 0 if sA == 0 else M * lda, B.data_ptr(), ldb, K * ldb, beta, C.data_ptr(), ldc, M * ldc, batch)
    ms = ev(f)
    ref = ev(lambda: torch.baddbmm(C, A.expand(batch, M, lda)[:, :, :K], B[:, :, :N], alpha=-1.0, beta=beta)) if M * N * batch < 2e9 else float("nan")
    print(f"gemm M={M:6d} N={N:6d} K={K:5d} b={batch:4d} beta={beta}: {ms:8.3f} ms  {2*M*N*K*batch/ms*1e-9:6.2f} TF/s   (cuBLAS {2*M*N*K*batch/ref*1e-9:6.2f})", flush=True)
def lu(n, batch, w):
    A = torch.randn(batch, n, n, dtype=torch.float64, device=dev) + n**0.5 * torch.eye(n, dtype=torch.float64, device=dev)
    R = torch.randn(batch, n, w, dtype=torch.float64, device=dev)
    need = ctypes.c_size_t(); lib.hps_lu_solve_workspace(batch, n, ctypes.byref(need))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev); info = torch.zeros(batch, dtype=torch.int32, device=dev)
    ptrs = (ctypes.c_void_p * 1)(R.data_ptr()); lds = (ctypes.c_int64 * 1)(w); st = (ctypes.c_int64 * 1)(n * w); nc = (ctypes.c_int * 1)(w)
    A0 = A.clone()
    def f():
        A.copy_(A0)
        lib.hps_lu_solve(_lib.stream_ptr(), batch, n, A.data_ptr(), n, n * n, 1, ptrs, lds, st, nc, ws.data_ptr(), ws.numel(), info.data_ptr())
    ms = ev(f, 2) - ev(lambda: A.copy_(A0), 2)
    fl = batch * (2 / 3 * n**3 + 2 * n * n * w)
    print(f"lu_solve n={n:6d} b={batch:4d} w={w:6d}: {ms:9.2f} ms  {fl/ms*1e-9:6.2f} TF/s", flush=True)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    gemm(15360, 38400, 128)            # root RHS trailing update
    gemm(15360, 15360, 128)            # root LU trailing update
    gemm(128, 38400, 128, beta=0.0)    # triangular multiply
    gemm(8192, 8192, 128); gemm(8192, 8192, 64); gemm(8192, 8192, 256); gemm(8192, 8192, 8192, beta=0.0)
    gemm(4000, 9600, 128, batch=8)     # level k=1
    gemm(872, 872, 128, batch=512); gemm(872, 600, 128, batch=512)   # leaf LU trailing / RHS
    gemm(968, 96, 32, batch=512)       # leaf inner-panel update
    gemm(1000, 600, 728, batch=512, beta=0.0, sA=None); gemm(600, 600, 1728, batch=512, beta=0.0, sA=0)
    gemm(100, 2400, 100, batch=64); gemm(400, 9600, 400, batch=8); gemm(1600, 38400, 1600, batch=1)
if which in ("all", "lu"):
    lu(1000, 512, 601); lu(1200, 64, 2401); lu(4800, 8, 9601); lu(19200, 1, 38401)
