"""Error study for an FP64-accurate GEMM built from INT8 tensor-core products (Ozaki scheme I),
on the CPU, BEFORE any kernel: how many 7-bit slices does the merge solve `D S = -C` need to stay inside the 1e-10 parity bar?

Each operand row (A) / column (B) is scaled by a power of two and cut into `s` signed slices of `w` bits; the slice products
A_i B_j are exact in int32 (K 2^(2w) < 2^31), and C = sum_{i+j<s} 2^(-w(i+j+2)) A_i B_j is accumulated in FP64.  That is what a
tcgen05 `kind::i8` kernel would compute; here the integer products are evaluated exactly in float64 (|values| < 2^53).

The operands are the real ones: one oct merge of a seeded 3D problem (p=8, q=6 -> D 432x432, C 432x864) is solved by a blocked
right-looking LU (partial pivoting inside 64-wide panels) in which EVERY block product (trailing update, forward and backward
substitution) goes through the emulated GEMM.  Reported: distance to the FP64 LAPACK solve (the parity metric), and distance of
both to a long-double-refined solution.  usage: python tools/ozaki_study.py"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _cases import hps, make_domain  # noqa: E402
from oracle import hps_oracle as orc  # noqa: E402

W = 7  # bits per slice: a signed 8-bit value


def split(M, axis, s):
    """Power-of-two scaling along `axis` and `s` integer slices of W bits (most significant first)."""
    mx = np.max(np.abs(M), axis=axis, keepdims=True)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)  # |M| 2^-e < 1/2... strictly below 1
    r = M * np.exp2(-e)
    out = []
    for _ in range(s):
        r = r * 2.0**W
        q = np.trunc(r)
        assert np.max(np.abs(q)) <= 2**W - 1
        out.append(q)
        r = r - q
    return out, e


def ozaki_gemm(A, B, s):
    K = A.shape[1]
    assert K * 2 ** (2 * W) < 2**31, "int32 accumulation would overflow"
    As, ea = split(A, 1, s)
    Bs, eb = split(B, 0, s)
    C = np.zeros((A.shape[0], B.shape[1]))
    for d in range(s - 1, -1, -1):  # small terms first
        G = np.zeros_like(C)
        for i in range(d + 1):
            G += As[i] @ Bs[d - i]  # exact: integers below 2^53
        C += G * 2.0 ** (-W * (d + 2))
    return C * np.exp2(ea) * np.exp2(eb)


def blocked_solve(D, R, gemm, nb=64):
    """X = D^-1 R through a blocked LU whose block products all go through `gemm`."""
    A = D.copy()
    X = R.copy()
    n = A.shape[0]
    for j in range(0, n, nb):
        e = min(n, j + nb)
        P, Lp, U = sla.lu(A[j:, j:e])  # panel with partial pivoting (FP64 on CUDA cores in the product path too)
        perm = np.argmax(P, axis=0)  # panel = P L U  ->  panel[perm] = L U
        A[j:] = A[j:][perm]
        X[j:] = X[j:][perm]
        blk = np.tril(Lp, -1)
        blk[: e - j] += U
        A[j:, j:e] = blk
        L11 = np.tril(A[j:e, j:e], -1) + np.eye(e - j)
        if e < n:
            A[j:e, e:] = sla.solve_triangular(L11, A[j:e, e:], lower=True, unit_diagonal=True)
            A[e:, e:] -= gemm(A[e:, j:e], A[j:e, e:])
        X[j:e] = sla.solve_triangular(L11, X[j:e], lower=True, unit_diagonal=True)
        if e < n:
            X[e:] -= gemm(A[e:, j:e], X[j:e])
    for j in range(((n - 1) // nb) * nb, -1, -nb):
        e = min(n, j + nb)
        X[j:e] = sla.solve_triangular(np.triu(A[j:e, j:e]), X[j:e], lower=False)
        if j > 0:
            X[:j] -= gemm(A[:j, j:e], X[j:e])
    return X


def merge_system(p=8, q=6, seed=3):
    rng = np.random.default_rng(seed)
    shp = (8, p**3)
    co = {f"{k}_coefficients": 1 + 0.1 * rng.normal(size=shp) for k in ("D_xx", "D_yy", "D_zz")}
    co["D_x_coefficients"] = rng.normal(size=shp)
    co["I_coefficients"] = rng.normal(size=shp)
    pb = hps.PDEProblem(make_domain(3, p, q, 1), source=rng.normal(size=shp), **co)
    _, T, _, _ = orc.local_solve_stage_uniform_3D_DtN(pb)
    m = T.shape[-1] // 6
    C = np.zeros((12 * m, 24 * m))
    D = np.zeros((12 * m, 12 * m))
    fs = lambda f: slice(f * m, (f + 1) * m)  # noqa: E731
    for c in range(8):  # same block placement as oracle.uniform_oct_merge_DtN
        roles = orc._OCT_ROLES[c]
        ext = [f for f in range(6) if roles[f][0] == "ext"]
        inn = [f for f in range(6) if roles[f][0] == "int"]
        for g in inn:
            for i, f in enumerate(ext):
                C[fs(roles[g][1]), fs(3 * c + i)] = T[c][fs(g), fs(f)]
            for f in inn:
                D[fs(roles[f][1]), fs(roles[g][1])] += T[c][fs(f), fs(g)]
    return D, -C


def refine_longdouble(D, R, X, iters=3):
    Dl, Rl, Xl = D.astype(np.longdouble), R.astype(np.longdouble), X.astype(np.longdouble)
    lu = sla.lu_factor(D)
    for _ in range(iters):
        res = Rl - Dl @ Xl
        Xl = Xl + sla.lu_solve(lu, res.astype(np.float64)).astype(np.longdouble)
    return Xl


def main():
    D, R = merge_system()
    print(f"D {D.shape}, rhs {R.shape}, cond(D) = {np.linalg.cond(D):.2e}, entry range of D: {np.abs(D[D != 0]).min():.1e} .. {np.abs(D).max():.1e}")
    S64 = sla.solve(D, R)
    Sx = refine_longdouble(D, R, S64)
    nrm = float(np.max(np.abs(Sx)))
    rel = lambda a, b: float(np.max(np.abs(a.astype(np.longdouble) - b)) / nrm)  # noqa: E731
    print(f"LAPACK FP64 vs long-double-refined:                {rel(S64, Sx):.2e}")
    Sb = blocked_solve(D, R, lambda a, b: a @ b)
    print(f"blocked LU, FP64 products vs refined:              {rel(Sb, Sx):.2e}   vs LAPACK {rel(Sb, S64):.2e}")
    for s in (4, 5, 6, 7, 8):
        t0 = time.time()
        So = blocked_solve(D, R, lambda a, b: ozaki_gemm(a, b, s))
        prods = s * (s + 1) // 2
        print(f"s={s} slices ({prods:2d} INT8 products, {W * s} bits): vs refined {rel(So, Sx):.2e}   vs LAPACK (parity metric) {rel(So, S64):.2e}"
              f"   [INT8-equivalent rate at the 4.5 POP/s nominal peak: {4500 / prods:.0f} TF/s]   ({time.time() - t0:.1f} s)")
    # one plain product with a wide dynamic range inside the rows, the known weak spot of a fixed-point split
    rng = np.random.default_rng(0)
    A = rng.normal(size=(256, 512)) * np.exp(4 * rng.normal(size=(256, 512)))
    B = rng.normal(size=(512, 256)) * np.exp(4 * rng.normal(size=(512, 256)))
    Cx = A.astype(np.longdouble) @ B.astype(np.longdouble)
    cw = lambda C: float(np.max(np.abs(C - Cx) / (np.abs(A).astype(np.longdouble) @ np.abs(B).astype(np.longdouble))))  # noqa: E731
    print(f"wide-range product (entries over ~14 decades), error relative to |A||B| element-wise: FP64 {cw(A @ B):.1e}, "
          + ", ".join(f"s={s}: {cw(ozaki_gemm(A, B, s)):.1e}" for s in (6, 8, 10)))


if __name__ == "__main__":
    main()
