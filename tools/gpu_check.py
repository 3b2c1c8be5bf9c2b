"""Developer smoke/parity sweep run on the GPU box (verbose; the committed tests live in tests/)."""
import ctypes, sys, time, traceback
import numpy as np, torch
sys.path.insert(0, ".")
from jaxhps_b200 import _lib
import jaxhps_b200 as hps
from oracle import hps_oracle as orc

lib = _lib.load()
dev = torch.device("cuda:0")
def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
results = []
def run(name, fn):
    t0 = time.time()
    try:
        r = fn(); torch.cuda.synchronize()
        print(f"[{'OK ' if r[0] else 'BAD'}] {name}: {r[1]}  ({time.time()-t0:.2f}s)", flush=True)
        results.append((name, r[0]))
    except Exception as e:
        traceback.print_exc()
        print(f"[EXC] {name}: {e}", flush=True)
        results.append((name, False))

def gemm_case(M, N, K, batch=1, alpha=1.0, beta=0.0, pad=0):
    def f():
        g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
        A = torch.randn(batch, M, K + pad, dtype=torch.float64, generator=g).to(dev)
        B = torch.randn(batch, K, N + pad, dtype=torch.float64, generator=g).to(dev)
        C = torch.randn(batch, M, N + pad, dtype=torch.float64, generator=g).to(dev)
        ref = alpha * torch.matmul(A[:, :, :K], B[:, :, :N]) + beta * C[:, :, :N]
        rc = lib.hps_dgemm_strided_batched(_lib.stream_ptr(), M, N, K, alpha, A.data_ptr(), K + pad, A.stride(0),
                                           B.data_ptr(), N + pad, B.stride(0), beta, C.data_ptr(), N + pad, C.stride(0), batch)
        _lib.check(rc, "gemm")
        e = rel(C[:, :, :N].cpu(), ref.cpu())
        return e < 1e-13, f"rel={e:.2e}"
    return f

def lu_case(n, batch, widths, cond_shift=0.0):
    def f():
        g = torch.Generator(device="cpu").manual_seed(n + batch)
        A = torch.randn(batch, n, n, dtype=torch.float64, generator=g).to(dev)
        if cond_shift: A += cond_shift * torch.eye(n, dtype=torch.float64, device=dev)
        rhs = [torch.randn(batch, n, w, dtype=torch.float64, generator=g).to(dev) for w in widths]
        ref = [torch.linalg.solve(A, r) for r in rhs]
        A2 = A.clone(); rhs2 = [r.clone() for r in rhs]
        need = ctypes.c_size_t(); lib.hps_lu_solve_workspace(batch, n, ctypes.byref(need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        info = torch.zeros(batch, dtype=torch.int32, device=dev)
        k = len(widths)
        ptrs = (ctypes.c_void_p * k)(*[r.data_ptr() for r in rhs2])
        lds = (ctypes.c_int64 * k)(*widths); strides = (ctypes.c_int64 * k)(*[n * w for w in widths]); nc = (ctypes.c_int * k)(*widths)
        torch.cuda.synchronize(); t0 = time.time()
        rc = lib.hps_lu_solve(_lib.stream_ptr(), batch, n, A2.data_ptr(), n, n * n, k, ptrs, lds, strides, nc, ws.data_ptr(), ws.numel(), info.data_ptr())
        _lib.check(rc, "lu"); torch.cuda.synchronize(); dt = time.time() - t0
        errs = [rel(a.cpu(), b.cpu()) for a, b in zip(rhs2, ref)]
        # residual check too
        res = max(float((torch.matmul(A, x) - r).abs().max() / r.abs().max()) for x, r in zip(rhs2, rhs))
        ok = int(info.abs().max()) == 0 and res < 1e-9
        return ok, f"rel_vs_torch={['%.1e' % e for e in errs]} resid={res:.1e} info={int(info.abs().max())} t={dt*1e3:.1f}ms"
    return f

def problem(p, q, L, dim=3, nsrc=1, seed=3):
    rng = np.random.default_rng(seed)
    if dim == 3:
        root = hps.DiscretizationNode3D(0., 1., 0., 1., 0., 1.)
    else:
        root = hps.DiscretizationNode2D(-1., 1., -1., 1.)
    dom = hps.Domain(p, q, root, L)
    shp = dom.interior_points[..., 0].shape
    co = {k: 1 + 0.1 * rng.normal(size=shp) for k in (("D_xx_coefficients", "D_yy_coefficients", "D_zz_coefficients") if dim == 3 else ("D_xx_coefficients", "D_yy_coefficients"))}
    co["D_xy_coefficients"] = 0.1 * rng.normal(size=shp)
    co["D_y_coefficients"] = rng.normal(size=shp)
    co["I_coefficients"] = rng.normal(size=shp)
    if dim == 3:
        co["D_z_coefficients"] = rng.normal(size=shp); co["D_yz_coefficients"] = 0.1 * rng.normal(size=shp)
    src = rng.normal(size=shp if nsrc == 1 else shp + (nsrc,))
    return hps.PDEProblem(dom, source=src, **co), rng

def stage_case(p, q, L, dim=3, nsrc=1):
    def f():
        pb, rng = problem(p, q, L, dim, nsrc)
        if dim == 3:
            o_ls, o_mg, o_dp = orc.local_solve_stage_uniform_3D_DtN, orc.merge_stage_uniform_3D_DtN, orc.down_pass_uniform_3D_DtN
            g_ls, g_mg, g_dp = hps.local_solve.local_solve_stage_uniform_3D_DtN, hps.merge.merge_stage_uniform_3D_DtN, hps.down_pass.down_pass_uniform_3D_DtN
        else:
            o_ls, o_mg, o_dp = orc.local_solve_stage_uniform_2D_DtN, orc.merge_stage_uniform_2D_DtN, orc.down_pass_uniform_2D_DtN
            g_ls, g_mg, g_dp = hps.local_solve.local_solve_stage_uniform_2D_DtN, hps.merge.merge_stage_uniform_2D_DtN, hps.down_pass.down_pass_uniform_2D_DtN
        Yo, To, vo, ho = o_ls(pb)
        Y, T, v, h = g_ls(pb)
        e_leaf = [rel(Y, Yo), rel(T, To), rel(v, vo), rel(h, ho)]
        So, go, Tt = o_mg(To, ho, L, return_T=True)
        S, g, Ttop = g_mg(To, ho, L, return_T=True)
        e_S = [rel(a, b) for a, b in zip(S, So)]; e_g = [rel(a, b) for a, b in zip(g, go)]; e_T = rel(Ttop, Tt)
        nb = pb.domain.boundary_points.shape[0]
        bd = rng.normal(size=(nb,) if nsrc == 1 else (nb, nsrc))
        uo = o_dp(bd, So, go, Yo, vo)
        u = g_dp(bd, So, go, Yo, vo)
        e_u = rel(u, uo)
        allv = e_leaf + e_S + e_g + [e_T, e_u]
        return max(allv) < 1e-10, f"leaf(Y,T,v,h)={['%.1e'%e for e in e_leaf]} S={['%.1e'%e for e in e_S]} g={['%.1e'%e for e in e_g]} Ttop={e_T:.1e} u={e_u:.1e}"
    return f

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    for (M, N, K, b) in ((128, 128, 16, 1), (128, 128, 64, 1), (100, 200, 50, 3), (1000, 600, 728, 2), (129, 131, 37, 2), (37, 16, 5, 1), (500, 17, 33, 4), (64, 4000, 128, 1)):
        run(f"gemm {M}x{N}x{K} b{b}", gemm_case(M, N, K, b))
    run("gemm alpha/beta", gemm_case(300, 260, 100, 2, alpha=-1.0, beta=1.0))
    run("gemm padded ld (odd)", gemm_case(200, 150, 64, 2, alpha=0.5, beta=2.0, pad=3))
    run("gemm skinny N=1", gemm_case(600, 1, 1728, 3))
    run("gemm skinny N=5", gemm_case(333, 5, 77, 3, alpha=2.0, beta=1.0))
if which in ("all", "lu"):
    for (n, b, w) in ((8, 2, [3]), (32, 3, [20]), (33, 2, [40, 1]), (100, 4, [64, 2]), (128, 2, [128]), (200, 3, [300, 1]),
                      (512, 2, [100]), (1000, 4, [600, 1]), (1200, 2, [2400, 1]), (2000, 2, [64]), (4800, 1, [200, 1]), (7000, 1, [128])):
        run(f"lu n={n} b={b} w={w}", lu_case(n, b, w))
if which in ("all", "stage"):
    run("3D p6 q4 L1", stage_case(6, 4, 1))
    run("3D p6 q4 L2", stage_case(6, 4, 2))
    run("3D p7 q5 L1 (odd)", stage_case(7, 5, 1))
    run("3D p5 q3 L2 nsrc3", stage_case(5, 3, 2, nsrc=3))
    run("3D p8 q6 L2", stage_case(8, 6, 2))
if which in ("all", "stage2d"):
    run("2D p8 q6 L2", stage_case(8, 6, 2, dim=2))
    run("2D p7 q5 L3 nsrc2", stage_case(7, 5, 3, dim=2, nsrc=2))
    run("2D p16 q14 L3", stage_case(16, 14, 3, dim=2))
bad = [n for n, ok in results if not ok]
print("SUMMARY:", len(results) - len(bad), "ok,", len(bad), "bad", bad)
