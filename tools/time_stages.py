"""Per-stage wall/CUDA timing of the 3D p=12 path (developer tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import jaxhps_b200 as hps
from jaxhps_b200.local_solve import local_solve_stage_uniform_3D_DtN
from jaxhps_b200.merge import merge_stage_uniform_3D_DtN
from jaxhps_b200.down_pass import down_pass_uniform_3D_DtN

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p, q = 12, 10
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
root = hps.DiscretizationNode3D(0., 1., 0., 1., 0., 1.)
t0 = time.time(); dom = hps.Domain(p, q, root, L); print("domain", time.time() - t0)
shp = dom.interior_points[..., 0].shape
c = torch.from_numpy(1 + 0.1 * rng.normal(size=shp)).to(dev)
src = torch.from_numpy(rng.normal(size=shp)).to(dev)
t0 = time.time(); pb = hps.PDEProblem(dom, source=src, D_xx_coefficients=c, D_yy_coefficients=c, D_zz_coefficients=c); print("pdeproblem", time.time() - t0)
g = torch.from_numpy(rng.normal(size=dom.boundary_points.shape[0])).to(dev)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    torch.cuda.synchronize()
    e0 = ev(); Y, T, v, h = local_solve_stage_uniform_3D_DtN(pb, device=dev, host_device=dev)
    e1 = ev(); S, gt = merge_stage_uniform_3D_DtN(T, h, L, device=dev, host_device=dev)
    e2 = ev(); u = down_pass_uniform_3D_DtN(g, S, gt, Y, v, device=dev, host_device=dev)
    e3 = ev(); torch.cuda.synchronize()
    n = shp[0]
    tl, tm, td = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
    print(f"iter {it}: local {tl:.1f} ms ({n/tl*1e3:.0f} leaves/s, {n*3.99e9/tl*1e-9:.2f} TF/s lean) merge {tm:.1f} ms down {td:.2f} ms total {tl+tm+td:.1f} ms  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GB")
    del Y, T, v, h, S, gt, u
