"""One large transposed mat-vec through the C ABI (ncu target / bandwidth check): S^T g at the L=3 root size."""
import sys
import torch
sys.path.insert(0, ".")
from jaxhps_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
M, K, N = 19200, 38400, 1
A = torch.randn(M, K, dtype=torch.float64, device=dev)
X = torch.randn(M, N, dtype=torch.float64, device=dev)
C = torch.empty(K, N, dtype=torch.float64, device=dev)
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.hps_gemv_t_strided_batched(_lib.stream_ptr(), M, K, N, 1.0, A.data_ptr(), K, 0, X.data_ptr(), N, 0, 0.0,
                                              C.data_ptr(), N, 0, 1, 0), "gemv_t")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"gemv_t {M}x{K}: {ms:.3f} ms  {M * K * 8 / ms * 1e-6:.0f} GB/s")
print("max err", float((C - A.T @ X).abs().max()))
