import sys, torch
sys.path.insert(0, ".")
from jaxhps_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
M, K = 19200, 38400
A = torch.randn(1, M, K, dtype=torch.float64, device=dev); x = torch.randn(1, K, 1, dtype=torch.float64, device=dev)
c = torch.randn(1, M, 1, dtype=torch.float64, device=dev); out = torch.empty_like(c)
for _ in range(3):
    lib.hps_leaf_apply(_lib.stream_ptr(), 1, M, K, 1, A.data_ptr(), x.data_ptr(), c.data_ptr(), out.data_ptr())
torch.cuda.synchronize()
