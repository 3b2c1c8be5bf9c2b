# Library comparators on the GPU box: cuBLAS DGEMM, cuSOLVER LU, PCIe copy. Dev tool.
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
out = {}
for n in (2048, 4096, 8192):
    A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = ev(lambda: torch.matmul(A, B))
    out[f"dgemm_{n}_tflops"] = 2*n**3/ms*1e-9
# sustained 3 s
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t0 = time.time(); cnt = 0
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
while time.time() - t0 < 3.0:
    for _ in range(4): torch.matmul(A, B); cnt += 1
    torch.cuda.synchronize()
b.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained_tflops"] = cnt*2*n**3/a.elapsed_time(b)*1e-9
# batched small gemm like leaf: 148 x (1000x1000x1000)
A = torch.randn(296, 1000, 1000, dtype=torch.float64, device=dev); B = torch.randn(296, 1000, 1000, dtype=torch.float64, device=dev)
ms = ev(lambda: torch.bmm(A, B)); out["bmm_296x1000_tflops"] = 296*2e9/ms*1e-9
# rank-64 update on 8192: C -= A(8192x64) B(64x8192)
C = torch.randn(8192, 8192, dtype=torch.float64, device=dev); A = torch.randn(8192, 64, dtype=torch.float64, device=dev); B = torch.randn(64, 8192, dtype=torch.float64, device=dev)
ms = ev(lambda: torch.addmm(C, A, B, alpha=-1, out=C)); out["rank64_update_8192_tflops"] = 2*8192*8192*64/ms*1e-9
A = torch.randn(8192, 128, dtype=torch.float64, device=dev); B = torch.randn(128, 8192, dtype=torch.float64, device=dev)
ms = ev(lambda: torch.addmm(C, A, B, alpha=-1, out=C)); out["rank128_update_8192_tflops"] = 2*8192*8192*128/ms*1e-9
del A, B, C
# LU
for n in (1000, 4800, 9600):
    bsz = {1000: 256, 4800: 8, 9600: 1}[n]
    M = torch.randn(bsz, n, n, dtype=torch.float64, device=dev) + n**0.5*torch.eye(n, dtype=torch.float64, device=dev)
    ms = ev(lambda: torch.linalg.lu_factor(M), n=2)
    out[f"cusolver_lu_{bsz}x{n}_tflops"] = bsz*(2/3)*n**3/ms*1e-9
    del M
# PCIe
h = torch.empty(1 << 28, dtype=torch.uint8).pin_memory(); d = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
ms = ev(lambda: d.copy_(h, non_blocking=True)); out["h2d_gbs"] = (1 << 28)/ms*1e-6
ms = ev(lambda: h.copy_(d, non_blocking=True)); out["d2h_gbs"] = (1 << 28)/ms*1e-6
x = torch.empty(1 << 30, dtype=torch.uint8, device=dev); y = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
ms = ev(lambda: y.copy_(x)); out["d2d_copy_gbs"] = 2*(1 << 30)/ms*1e-6
import os; out["nproc"] = os.cpu_count()
print(json.dumps(out, indent=1))
