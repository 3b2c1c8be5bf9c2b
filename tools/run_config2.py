"""BASELINE config 2: 2D variable-coefficient Helmholtz, ItI, p=16 q=14, k=100, complex128 (developer run).
 (a) constant-coefficient plane wave at L=6: error vs the analytic solution;
 (b) gauss-bump potential at L=6: timing; at L=4: parity vs the oracle."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import jaxhps_b200 as hps
from oracle import hps_oracle as orc

k = 100.0
def problem(L, bumps):
    dom = hps.Domain(16, 14, hps.DiscretizationNode2D(-1., 1., -1., 1.), L)
    x = dom.interior_points
    one = np.ones_like(x[..., 0])
    if bumps:
        rng = np.random.default_rng(0)
        centres = rng.uniform(-0.5, 0.5, size=(10, 2))
        q = sum(np.exp(-50 * ((x[..., 0] - c[0]) ** 2 + (x[..., 1] - c[1]) ** 2)) for c in centres)
        I = k**2 * (1 + q)
        src = -k**2 * q * np.exp(1j * k * x[..., 0])
    else:
        I = k**2 * one
        src = np.zeros_like(one, dtype=np.complex128)
    pb = hps.PDEProblem(dom, source=src, D_xx_coefficients=one, D_yy_coefficients=one, I_coefficients=I, use_ItI=True, eta=k)
    b = dom.boundary_points; n = b.shape[0] // 4
    ub = np.exp(1j * k * b[:, 0]); nx = np.concatenate([np.zeros(n), np.ones(n), np.zeros(n), -np.ones(n)])
    g = nx * 1j * k * ub + 1j * k * ub
    return dom, pb, g

def timed(pb, g):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hps.build_solver(pb, host_device="cuda:0"); torch.cuda.synchronize(); t1 = time.perf_counter()
    u = hps.solve(pb, g); torch.cuda.synchronize(); t2 = time.perf_counter()
    return u, t1 - t0, t2 - t1

dom, pb, g = problem(6, False)
timed(pb, g); pb.reset()
u, tb, ts = timed(pb, g)
ex = np.exp(1j * k * dom.interior_points[..., 0])
print(f"(a) plane wave L=6 (4096 leaves, k=100): build {tb*1e3:.1f} ms solve {ts*1e3:.1f} ms  max|u-exact| = {np.abs(u-ex).max():.2e}", flush=True)
dom, pb, g = problem(6, True)
timed(pb, g); pb.reset()
u, tb, ts = timed(pb, g)
print(f"(b) gauss-bump potential L=6: build {tb*1e3:.1f} ms solve {ts*1e3:.1f} ms  ({4096/(tb+ts):.0f} leaves/s)  |u|max={np.abs(u).max():.3f}", flush=True)
dom, pb, g = problem(4, True)
hps.build_solver(pb); u = hps.solve(pb, g)
t0 = time.perf_counter()
Y, R, v, h = orc.local_solve_stage_uniform_2D_ItI(pb); S, gt = orc.merge_stage_uniform_2D_ItI(R, h, 4); uo = orc.down_pass_uniform_2D_ItI(g, S, gt, Y, v)
print(f"    L=4 parity vs oracle: rel err {np.abs(u-uo).max()/np.abs(uo).max():.2e}  (oracle {time.perf_counter()-t0:.1f} s on CPU for 256 leaves)", flush=True)
