"""Extended-precision (x87 80-bit ``numpy.longdouble``, eps = 1.1e-19) evaluation of the ItI leaf operators, used to
ARBITRATE between the CUDA path and the FP64 oracle where the two legitimately differ by more than 1e-10: the ItI leaf
systems are ill-conditioned (cond ~ 2e5 at p=16), so two correct FP64 evaluations can sit 1e-11..1e-9 apart.  "Truth"
is the exact-arithmetic result for the FP64 INPUTS the reference defines (coefficient fields, first-derivative
matrices D_x, D_y, P, QH, G), with the second derivatives formed as D_x D_x, D_y D_y, D_x D_y in extended precision
(reference `_precompute_operators_2D.py:45-59`, `local_solve/_uniform_2D_ItI.py:120-186`)."""
import numpy as np

LD = np.longdouble
CLD = np.clongdouble


def lu_solve_ld(B, R):
    """Solve B X = R with partial pivoting, everything in complex extended precision."""
    A = np.array(B, dtype=CLD)
    X = np.array(R, dtype=CLD)
    n = A.shape[0]
    for k in range(n):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        if p != k:
            A[[k, p]] = A[[p, k]]
            X[[k, p]] = X[[p, k]]
        l = A[k + 1:, k] / A[k, k]
        A[k + 1:, k + 1:] -= np.outer(l, A[k, k + 1:])
        X[k + 1:] -= np.outer(l, X[k])
    for k in range(n - 1, -1, -1):
        X[k] = (X[k] - A[k, k + 1:] @ X[k + 1:]) / A[k, k]
    return X


def iti_leaf_truth(pb, leaf, rounded=True):
    """(Y, R, v, h) of one leaf in extended precision (rounded to complex128 on return unless ``rounded=False``)."""
    Dx, Dy = np.asarray(pb.D_x, dtype=LD), np.asarray(pb.D_y, dtype=LD)
    ops = {"D_xx": Dx @ Dx, "D_xy": Dx @ Dy, "D_yy": Dy @ Dy, "D_x": Dx, "D_y": Dy, "I": np.eye(Dx.shape[0], dtype=LD)}
    n_c = Dx.shape[0]
    A = np.zeros((n_c, n_c), dtype=CLD)
    for name, D in ops.items():
        c = getattr(pb, f"{name}_coefficients", None)
        if c is not None:
            A = A + np.asarray(c[leaf], dtype=CLD)[:, None] * D
    P, QH, G = (np.asarray(x, dtype=CLD) for x in (pb.P, pb.QH, pb.G))
    nb = P.shape[0]
    src = np.asarray(pb.source)[leaf]
    src = src[:, None] if src.ndim == 1 else src
    B = np.concatenate([G, A[nb:]], axis=0)
    rhs = np.zeros((n_c, P.shape[1] + src.shape[1]), dtype=CLD)
    rhs[:nb, : P.shape[1]] = P
    rhs[nb:, P.shape[1]:] = src[nb:]
    X = lu_solve_ld(B, rhs)
    Y, v = X[:, : P.shape[1]], X[:, P.shape[1]:]
    R, h = QH @ Y, QH @ v
    if not rounded:
        return Y, R, v, h
    return tuple(np.asarray(t, dtype=np.complex128) for t in (Y, R, v, h))


def iti_pipeline_truth(pb, L, boundary_data):
    """The whole 2D ItI build + solve in extended precision: leaf solves above, then the ORACLE's own merge and down-pass
    code evaluated on ``clongdouble`` arrays (oracle/hps_oracle.py dispatches its inverses to extended-precision
    Gaussian elimination for such inputs).  Returns ``(Y, R, v, h, S_lst, g_lst, R_top, u)`` rounded to complex128;
    single-source only.  Cost: ~0.2 s per p=16 leaf, seconds per merge level — meant for L <= 3."""
    from oracle import hps_oracle as orc

    n_leaves = 4**L
    leaves = [iti_leaf_truth(pb, i, rounded=False) for i in range(n_leaves)]
    Y, R, v, h = (np.stack([lf[k] for lf in leaves]) for k in range(4))
    S_lst, g_lst, R_top = orc.merge_stage_uniform_2D_ItI(R, h[..., 0], L, return_T=True)
    u = orc.down_pass_uniform_2D_ItI(np.asarray(boundary_data, dtype=CLD), S_lst, g_lst, Y, v[..., 0])
    c = lambda a: np.asarray(a, dtype=np.complex128)  # noqa: E731
    return c(Y), c(R), c(v[..., 0]), c(h[..., 0]), [c(S) for S in S_lst], [c(g) for g in g_lst], c(R_top), c(u)
