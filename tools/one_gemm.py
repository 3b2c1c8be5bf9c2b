import sys, torch
sys.path.insert(0, ".")
from jaxhps_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
M, N, K = (int(x) for x in sys.argv[1:4]); beta = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(K, N, dtype=torch.float64, device=dev); C = torch.randn(M, N, dtype=torch.float64, device=dev)
for _ in range(3):
    lib.hps_dgemm_strided_batched(_lib.stream_ptr(), M, N, K, -1.0, A.data_ptr(), K, 0, B.data_ptr(), N, 0, beta, C.data_ptr(), N, 0, 1)
torch.cuda.synchronize()
