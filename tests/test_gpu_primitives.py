"""GPU parity of the dense building blocks, called through the C ABI (ctypes)."""
import ctypes

import numpy as np
import pytest
import torch

from jaxhps_b200 import _lib

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize(
    "M,N,K,batch,alpha,beta,pad",
    [
        (128, 128, 16, 1, 1.0, 0.0, 0),
        (1000, 600, 728, 2, -1.0, 0.0, 0),  # leaf-shaped: A_ie P
        (129, 131, 37, 2, 1.0, 1.0, 0),  # ragged edges
        (200, 150, 64, 2, 0.5, 2.0, 3),  # odd leading dimensions -> 8-byte copy path
        (37, 16, 5, 1, 1.0, 0.0, 0),
        (600, 1, 1728, 3, 1.0, 0.0, 0),  # narrow kernel (h = Q v)
        (333, 5, 77, 3, 2.0, 1.0, 0),
        (0, 10, 10, 1, 1.0, 0.0, 0),  # empty
    ],
)
def test_dgemm_matches_fp64_reference(M, N, K, batch, alpha, beta, pad):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(batch, M, K + pad, dtype=torch.float64, generator=g).to(dev)
    B = torch.randn(batch, K, N + pad, dtype=torch.float64, generator=g).to(dev)
    C = torch.randn(batch, M, N + pad, dtype=torch.float64, generator=g).to(dev)
    ref = alpha * torch.matmul(A[:, :, :K], B[:, :, :N]) + beta * C[:, :, :N]
    rc = lib.hps_dgemm_strided_batched(_lib.stream_ptr(), M, N, K, alpha, A.data_ptr(), K + pad, A.stride(0),
                                       B.data_ptr(), N + pad, B.stride(0), beta, C.data_ptr(), N + pad, C.stride(0), batch)
    _lib.check(rc, "hps_dgemm_strided_batched")
    torch.cuda.synchronize()
    if M:
        assert _rel(C[:, :, :N], ref) < 1e-13


def _lu_solve(A, rhs):
    lib = _lib.load()
    batch, n, _ = A.shape
    need = ctypes.c_size_t()
    lib.hps_lu_solve_workspace(batch, n, ctypes.byref(need))
    ws = torch.empty(need.value, dtype=torch.uint8, device=A.device)
    info = torch.zeros(batch, dtype=torch.int32, device=A.device)
    k = len(rhs)
    widths = [r.shape[-1] for r in rhs]
    ptrs = (ctypes.c_void_p * k)(*[r.data_ptr() for r in rhs])
    lds = (ctypes.c_int64 * k)(*widths)
    strides = (ctypes.c_int64 * k)(*[n * w for w in widths])
    nc = (ctypes.c_int * k)(*widths)
    rc = lib.hps_lu_solve(_lib.stream_ptr(), batch, n, A.data_ptr(), n, n * n, k, ptrs, lds, strides, nc,
                          ws.data_ptr(), ws.numel(), info.data_ptr())
    _lib.check(rc, "hps_lu_solve")
    torch.cuda.synchronize()
    return info


@pytest.mark.parametrize(
    "n,batch,widths",
    [
        (8, 2, [3]),
        (33, 2, [40, 1]),  # ragged inner panel + narrow rhs
        (128, 2, [128]),
        (200, 3, [300, 1]),
        (196, 5, [57]),  # largest block column one CTA holds in shared memory
        (197, 2, [10]),  # smallest 2-CTA cluster (128 + 69 rows)
        (777, 2, [5]),  # ragged last block column, 4-CTA cluster
        (1000, 3, [600, 1]),  # leaf size: 6-CTA cluster block column
        (1200, 2, [2400, 1]),  # first merge level size
        (1700, 2, [16]),  # 9 CTAs per matrix: smallest cooperative launch
        (2500, 20, [8]),  # cooperative launch split over several batches of matrices
        (4800, 1, [96]),  # 25 co-resident CTAs
        (7000, 1, [64]),
    ],
)
def test_lu_solve_residual_and_agreement(n, batch, widths):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n + batch)
    # well-conditioned, NOT diagonally dominant: pivoting is exercised
    A = torch.randn(batch, n, n, dtype=torch.float64, generator=g).to(dev)
    Q, _ = torch.linalg.qr(A)
    s = torch.linspace(1.0, 50.0, n, dtype=torch.float64, device=dev)
    A = (Q * s) @ torch.roll(Q, 1, dims=1).transpose(1, 2)
    rhs = [torch.randn(batch, n, w, dtype=torch.float64, generator=g).to(dev) for w in widths]
    ref = [torch.linalg.solve(A, r) for r in rhs]
    A2, rhs2 = A.clone(), [r.clone() for r in rhs]
    info = _lu_solve(A2, rhs2)
    assert int(info.abs().max()) == 0
    for x, r, b in zip(rhs2, ref, rhs):
        assert _rel(x, r) < 1e-10
        assert float((torch.matmul(A, x) - b).abs().max() / b.abs().max()) < 1e-11


def test_lu_reports_exact_singularity():
    dev = torch.device("cuda:0")
    A = torch.eye(40, dtype=torch.float64, device=dev).repeat(2, 1, 1)
    A[1, 17, 17] = 0.0
    b = torch.ones(2, 40, 1, dtype=torch.float64, device=dev)
    info = _lu_solve(A, [b])
    assert info.tolist() == [0, 18]


@pytest.mark.parametrize(
    "M,K,N,batch,alpha,beta,cplx",
    [
        (256, 56, 1, 64, 1.0, 0.0, False),     # Y^T w on many leaves
        (56, 256, 2, 64, 1.0, 0.0, True),      # QH^T h (complex, shared operator is covered by the adjoint tests)
        (2000, 300, 3, 1, -1.0, 0.0, False),   # few CTAs: the rows are split and summed atomically
        (2000, 300, 3, 1, 0.5, 2.0, True),     # ... with beta and complex entries
        (777, 129, 11, 2, 1.0, 1.0, False),    # more than 8 right-hand sides: two column groups
        (0, 5, 1, 1, 1.0, 0.0, False),         # empty
    ],
)
def test_transposed_matvec_matches_fp64_reference(M, K, N, batch, alpha, beta, cplx):
    """``hps_gemv_t_strided_batched``: C = alpha A^T X + beta C (plain transpose for complex128), through the raw C ABI."""
    lib = _lib.load()
    dev = torch.device("cuda:0")
    dt = torch.complex128 if cplx else torch.float64
    g = torch.Generator().manual_seed(M + 3 * K + 7 * N)
    A = torch.randn(batch, M, K, dtype=dt, generator=g).to(dev)
    X = torch.randn(batch, M, N, dtype=dt, generator=g).to(dev)
    C = torch.randn(batch, K, N, dtype=dt, generator=g).to(dev)
    ref = alpha * torch.matmul(A.transpose(1, 2), X) + beta * C
    rc = lib.hps_gemv_t_strided_batched(_lib.stream_ptr(), M, K, N, alpha, A.data_ptr(), K, M * K, X.data_ptr(), N, M * N, beta,
                                        C.data_ptr(), N, K * N, batch, 1 if cplx else 0)
    _lib.check(rc, "hps_gemv_t_strided_batched")
    torch.cuda.synchronize()
    if M:
        assert _rel(C, ref) < 1e-13
