"""LU accuracy study (developer tool, GPU): the embedded ItI leaf system of config-2-like leaves through hps_lu_solve
versus LAPACK and an extended-precision solve.  Reports forward error against the extended-precision solution,
normwise backward error, max |L| (must be <= 1 under partial pivoting) and the first diagonal entry of U that differs
from LAPACK's (a different pivot choice)."""
import ctypes
import sys

import numpy as np
import scipy.linalg as sla
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from _cases import rel_err, seeded_problem  # noqa: E402
from _longdouble import lu_solve_ld  # noqa: E402
from jaxhps_b200 import _lib  # noqa: E402
from oracle import hps_oracle as orc  # noqa: E402


def gpu_lu_solve(A, rhs):
    lib = _lib.load()
    batch, n, _ = A.shape
    need = ctypes.c_size_t()
    lib.hps_lu_solve_workspace(batch, n, ctypes.byref(need))
    ws = torch.empty(need.value, dtype=torch.uint8, device=A.device)
    info = torch.zeros(batch, dtype=torch.int32, device=A.device)
    w = rhs.shape[-1]
    ptrs = (ctypes.c_void_p * 1)(rhs.data_ptr())
    lds = (ctypes.c_int64 * 1)(w)
    strides = (ctypes.c_int64 * 1)(n * w)
    nc = (ctypes.c_int * 1)(w)
    rc = lib.hps_lu_solve(_lib.stream_ptr(), batch, n, A.data_ptr(), n, n * n, 1, ptrs, lds, strides, nc,
                          ws.data_ptr(), ws.numel(), info.data_ptr())
    _lib.check(rc, "hps_lu_solve")
    torch.cuda.synchronize()
    return info


def study(name, Be, re, batches=(1, 64)):
    n = Be.shape[0]
    Xt = np.asarray(lu_solve_ld(Be, re).real, dtype=np.float64)
    Xn = np.linalg.solve(Be, re)
    lu, piv = sla.lu_factor(Be)
    print(f"{name}: n={n} cond={np.linalg.cond(Be):.2e}  LAPACK vs truth {rel_err(Xn, Xt):.2e}")
    dev = torch.device("cuda:0")
    for batch in batches:
        A = torch.from_numpy(Be).to(dev).unsqueeze(0).repeat(batch, 1, 1).contiguous()
        R = torch.from_numpy(re).to(dev).unsqueeze(0).repeat(batch, 1, 1).contiguous()
        gpu_lu_solve(A, R)
        for b in sorted({0, batch - 1}):
            X = R[b].cpu().numpy()
            F = A[b].cpu().numpy()
            bwd = np.abs(Be @ X - re).max() / (np.abs(Be).sum(1).max() * np.abs(X).max())
            bwd_n = np.abs(Be @ Xn - re).max() / (np.abs(Be).sum(1).max() * np.abs(Xn).max())
            dU, dL = np.abs(np.diag(F)), np.abs(np.diag(lu))
            diff = np.nonzero(np.abs(dU - dL) > 1e-8 * dL)[0]
            print(f"   batch={batch} mat {b}: GPU vs truth {rel_err(X, Xt):.2e}  backward {bwd:.2e} (LAPACK {bwd_n:.2e})  "
                  f"max|L| {np.abs(np.tril(F, -1)).max():.6f}  max|U| {np.abs(np.triu(F)).max():.3e} (LAPACK {np.abs(np.triu(lu)).max():.3e})  "
                  f"first differing pivot {int(diff[0]) if diff.size else None} of {diff.size}")


def main():
    p, q, L = 16, 14, 3
    pb, _ = seeded_problem(20, p, q, L, 1, seed=100 + p)
    coeffs, which = orc.gather_coeffs(pb, orc._COEFF_ORDER_2D)
    ops = orc._leaf_operators(pb, orc._COEFF_ORDER_2D)
    A = orc.assemble_diff_operator(coeffs[:, 0], which, ops)
    nb = pb.P.shape[0]
    n = A.shape[0]
    B = np.concatenate([pb.G, A[nb:].astype(np.complex128)], 0)
    rhs = np.zeros((n, pb.P.shape[1]), complex)
    rhs[:nb] = pb.P
    Be = np.block([[B.real, -B.imag], [B.imag, B.real]])
    re = np.concatenate([rhs.real, rhs.imag], 0)
    study("ItI leaf p=16 (embedded)", Be, re)
    rng = np.random.default_rng(0)
    for n in (128, 196, 197, 512, 777):
        Q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
        Q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
        M = (Q1 * np.logspace(0, -5.3, n)) @ Q2.T
        study(f"random cond 2e5 n={n}", M, rng.standard_normal((n, 8)), batches=(1, 160))


if __name__ == "__main__":
    main()
