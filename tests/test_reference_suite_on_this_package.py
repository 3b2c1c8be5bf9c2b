"""Drop-in check: the REFERENCE's own test files for the host layer, run unmodified against this package
through a module alias (`tests/reference_alias/jaxhps` maps `jaxhps.*` onto `jaxhps_b200.*`; `jax` is the NumPy
shim of tests/golden).  Needs /root/reference (present where the CPU suite runs; skipped elsewhere)."""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_TESTS = "/root/reference/tests"

FILES = [
    "test_domain.py",  # Domain: uniform/adaptive construction, interpolation to/from regular grids
    "test_pdeproblem.py",  # PDEProblem validation, operators, chunking
    "test_precompute_operators_2D.py",  # P, Q, N, G, QH, projection operators
    "test_discretization_tree_operations_2D.py",  # splitting, per-side counts, corner search, paths
    "test_discretization_tree_operations_3D.py",
    "test_adaptive_discretization_2D.py",  # level-restricted mesh generation, L2 norms
    "test_adaptive_discretization_3D.py",
    "test_grid_creation_2D.py",  # point clouds, subdivision (through reference-signature adapters in the alias)
    "test_grid_creation_3D.py",
    "test_precompute_operators_3D.py",  # differentiation operators, P, Q, projection / refinement operators
    "test_interpolation_methods.py",  # HPS grid <-> regular grid
    "test_quadrature",  # the whole directory: points, weights, differentiation and interpolation matrices
]


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not available here")
def test_reference_host_tests_pass_against_this_package(tmp_path):
    dst = tmp_path / "tests"
    dst.mkdir()
    for f in FILES:
        src = os.path.join(REF_TESTS, f)
        if os.path.isdir(src):
            shutil.copytree(src, dst / f)
        else:
            shutil.copy(src, dst / f)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "reference_alias"), ROOT, os.path.join(HERE, "golden", "jaxshim")])
    out = subprocess.run([sys.executable, "-m", "pytest", str(dst), "-q", "-p", "no:cacheprovider"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=900)
    tail = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:]
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert " passed" in tail and "failed" not in tail, tail
    n_passed = int(tail.split(" passed")[0].split()[-1])
    assert n_passed >= 120, tail


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not available here")
def test_reference_accuracy_suite_passes_on_the_oracle(tmp_path):
    """The reference's own analytic known-answer suite (`tests/test_accuracy`: leaf DtN / ItI maps, single merges,
    source-free build + up pass) executed unmodified with the CPU ORACLE as the stage functions
    (`tests/reference_alias_oracle`): the strongest pin of the oracle SURVEY section 8(c) asks for."""
    dst = tmp_path / "tests"
    shutil.copytree(os.path.join(REF_TESTS, "test_accuracy"), dst / "test_accuracy")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "reference_alias_oracle"), ROOT, os.path.join(HERE, "golden", "jaxshim")])
    out = subprocess.run([sys.executable, "-m", "pytest", str(dst), "-q", "-p", "no:cacheprovider"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=900)
    tail = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:]
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert "failed" not in tail and int(tail.split(" passed")[0].split()[-1]) >= 14, tail


_MESH_SCRIPT = r"""
import sys, json
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, "/root/reference/src")
import jax.numpy as jnp
import jaxhps as ref
from jaxhps._discretization_tree import get_all_leaves
out = {}
for case in json.loads(sys.argv[2]):
    dim, p, q, tol, l2, seed = case
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.6, 0.6, size=(3, dim)); w = rng.uniform(10, 60, size=3); a = rng.uniform(0.5, 1.5, size=3)
    f = lambda x: sum(a[k] * jnp.exp(-w[k] * sum((x[..., d] - c[k, d]) ** 2 for d in range(dim))) for k in range(3))
    root = (ref.DiscretizationNode2D(xmin=-1., xmax=1., ymin=-1., ymax=1.) if dim == 2 else
            ref.DiscretizationNode3D(xmin=-1., xmax=1., ymin=-1., ymax=1., zmin=-1., zmax=1.))
    ref.Domain.from_adaptive_discretization(p=p, q=q, root=root, f=f, tol=tol, use_l_2_norm=bool(l2))
    keys = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[: 2 * dim]
    out[str(seed)] = [[float(getattr(l, k)) for k in keys] for l in get_all_leaves(root)]
print(json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not available here")
def test_mesh_generator_reproduces_the_reference_on_random_functions():
    """Level-restricted adaptive meshes for sums of random Gaussian bumps (2D / 3D, L_inf / L_2 criterion): this
    package's generator must produce exactly the leaves the reference's generator produces (run on the NumPy shim)."""
    import json

    import numpy as np

    import jaxhps_b200 as hps
    from jaxhps_b200._tree import get_all_leaves

    cases = [(2, 8, 6, 1e-3, 0, 101), (2, 6, 4, 1e-2, 1, 102), (2, 10, 8, 1e-5, 0, 103), (3, 4, 2, 3e-2, 0, 104),
             (3, 6, 4, 3e-2, 1, 105), (3, 5, 4, 1e-2, 0, 106)]
    shim = os.path.join(HERE, "golden", "jaxshim")
    out = subprocess.run([sys.executable, "-c", _MESH_SCRIPT, shim, json.dumps(cases)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    ref_leaves = json.loads(out.stdout.strip().splitlines()[-1])
    n_total = 0
    for dim, p, q, tol, l2, seed in cases:
        rng = np.random.default_rng(seed)
        c, w, a = rng.uniform(-0.6, 0.6, size=(3, dim)), rng.uniform(10, 60, size=3), rng.uniform(0.5, 1.5, size=3)

        def f(x, c=c, w=w, a=a, dim=dim):
            return sum(a[k] * np.exp(-w[k] * sum((x[..., d] - c[k, d]) ** 2 for d in range(dim))) for k in range(3))

        root = (hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0) if dim == 2 else hps.DiscretizationNode3D(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0))
        hps.Domain.from_adaptive_discretization(p=p, q=q, root=root, f=f, tol=tol, use_l_2_norm=bool(l2))
        keys = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[: 2 * dim]
        mine = [[float(getattr(leaf, k)) for k in keys] for leaf in get_all_leaves(root)]
        assert mine == ref_leaves[str(seed)], (dim, p, q, tol, l2, len(mine), len(ref_leaves[str(seed)]))
        n_total += len(mine)
    assert n_total > 100


_OPS_SCRIPT = r"""
import sys, json
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, "/root/reference/src")
import jax.numpy as jnp
import jaxhps as ref
np.savez(sys.argv[3], **{})
out = {}
for dim, p, q, L, iti in json.loads(sys.argv[2]):
    root = (ref.DiscretizationNode2D(xmin=-1., xmax=2., ymin=0., ymax=3.) if dim == 2 else
            ref.DiscretizationNode3D(xmin=-1., xmax=1., ymin=0., ymax=2., zmin=1., zmax=3.))
    dom = ref.Domain(p=p, q=q, root=root, L=L)
    z = jnp.zeros(dom.interior_points[..., 0].shape)
    pb = ref.PDEProblem(dom, source=z, D_xx_coefficients=z, **(dict(use_ItI=True, eta=2.5) if iti else {}))
    tag = f"{dim}_{p}_{q}_{L}_{int(iti)}"
    for name in ("P", "Q", "D_x", "D_y", "D_xx", "D_xy", "D_yy", "D_z", "D_zz", "D_xz", "D_yz", "G", "QH"):
        v = getattr(pb, name, None)
        if v is not None:
            out[tag + ":" + name] = np.asarray(v)
    out[tag + ":interior"] = np.asarray(dom.interior_points)
    out[tag + ":boundary"] = np.asarray(dom.boundary_points)
np.savez(sys.argv[3], **out)
"""


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not available here")
def test_precomputed_operators_and_point_clouds_match_the_reference_live(tmp_path):
    """Differential check over a range of orders: every pre-computed operator and both point clouds of the
    reference's PDEProblem / Domain (run on the NumPy shim) against this package's host layer."""
    import json

    import numpy as np

    import jaxhps_b200 as hps

    cases = [(2, 5, 3, 1, 0), (2, 8, 6, 2, 0), (2, 9, 7, 1, 1), (2, 12, 10, 1, 1), (3, 4, 2, 1, 0), (3, 5, 4, 1, 0), (3, 7, 5, 1, 0)]
    shim = os.path.join(HERE, "golden", "jaxshim")
    npz = str(tmp_path / "ref_ops.npz")
    out = subprocess.run([sys.executable, "-c", _OPS_SCRIPT, shim, json.dumps(cases), npz], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    R = np.load(npz)
    checked = 0
    for dim, p, q, L, iti in cases:
        root = (hps.DiscretizationNode2D(-1.0, 2.0, 0.0, 3.0) if dim == 2 else hps.DiscretizationNode3D(-1.0, 1.0, 0.0, 2.0, 1.0, 3.0))
        dom = hps.Domain(p, q, root, L)
        z = np.zeros(dom.interior_points[..., 0].shape)
        pb = hps.PDEProblem(dom, source=z, D_xx_coefficients=z, **(dict(use_ItI=True, eta=2.5) if iti else {}))
        tag = f"{dim}_{p}_{q}_{L}_{int(iti)}"
        for key in [k for k in R.files if k.startswith(tag + ":")]:
            name = key.split(":")[1]
            mine = {"interior": dom.interior_points, "boundary": dom.boundary_points}.get(name)
            if mine is None:
                mine = getattr(pb, name)
            ref_v = R[key]
            scale = max(1.0, float(np.abs(ref_v).max()))
            assert mine.shape == ref_v.shape, key
            assert np.abs(mine - ref_v).max() <= 1e-12 * scale, (key, float(np.abs(mine - ref_v).max()))
            checked += 1
    assert checked >= 50


_STAGE_SCRIPT = r"""
import sys, json
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, "/root/reference/src"); sys.path.insert(0, sys.argv[4])
import jax.numpy as jnp
import jaxhps as ref
from jaxhps.local_solve import local_solve_stage_uniform_2D_DtN, local_solve_stage_uniform_2D_ItI, local_solve_stage_uniform_3D_DtN
from jaxhps.merge import merge_stage_uniform_2D_DtN, merge_stage_uniform_2D_ItI, merge_stage_uniform_3D_DtN
from jaxhps.down_pass import down_pass_uniform_2D_DtN, down_pass_uniform_2D_ItI, down_pass_uniform_3D_DtN
from _cases import seeded_inputs, seeded_inputs_iti, ETA
out = {}
for dim, p, q, L, nsrc, seed in json.loads(sys.argv[2]):
    iti = dim == 20
    co, src, bdry = seeded_inputs_iti(p, q, L, nsrc, seed) if iti else seeded_inputs(dim, p, q, L, nsrc, seed)
    if dim == 3:
        root = ref.DiscretizationNode3D(xmin=0., xmax=1., ymin=0., ymax=1., zmin=0., zmax=1.)
        ls, mg, dp = local_solve_stage_uniform_3D_DtN, merge_stage_uniform_3D_DtN, down_pass_uniform_3D_DtN
    else:
        root = ref.DiscretizationNode2D(xmin=-1., xmax=1., ymin=-1., ymax=1.)
        ls, mg, dp = ((local_solve_stage_uniform_2D_ItI, merge_stage_uniform_2D_ItI, down_pass_uniform_2D_ItI) if iti else
                      (local_solve_stage_uniform_2D_DtN, merge_stage_uniform_2D_DtN, down_pass_uniform_2D_DtN))
    dom = ref.Domain(p=p, q=q, root=root, L=L)
    pb = ref.PDEProblem(dom, source=jnp.array(src), **{k: jnp.array(v) for k, v in co.items()}, **(dict(use_ItI=True, eta=ETA) if iti else {}))
    Y, T, v, h = ls(pb)
    S_lst, g_lst, T_top = mg(T, h, l=L, return_T=True)
    u = dp(jnp.array(bdry), S_lst, g_lst, Y, v)
    tag = f"{dim}_{p}_{q}_{L}_{nsrc}_{seed}"
    out[tag + ":u"], out[tag + ":T_top"], out[tag + ":h"] = np.asarray(u), np.asarray(T_top), np.asarray(h)
    out[tag + ":g_root"] = np.asarray(g_lst[-1])
np.savez(sys.argv[3], **out)
"""


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not available here")
def test_oracle_matches_the_reference_live_on_more_seeded_problems(tmp_path):
    """Beyond the committed fixtures: the reference's uniform stage functions (on the NumPy shim) and the oracle on six
    further seeded problems (2D DtN, 2D ItI, 3D DtN; single and multi-source)."""
    import json

    import numpy as np

    from oracle import hps_oracle as orc
    from _cases import rel_err, seeded_problem

    cases = [(2, 9, 7, 2, 1, 201), (2, 6, 4, 3, 2, 202), (20, 7, 5, 2, 1, 203), (20, 6, 4, 1, 3, 204), (3, 5, 3, 2, 1, 205),
             (3, 6, 4, 1, 1, 206)]
    shim = os.path.join(HERE, "golden", "jaxshim")
    npz = str(tmp_path / "ref_stage.npz")
    out = subprocess.run([sys.executable, "-c", _STAGE_SCRIPT, shim, json.dumps(cases), npz, HERE], capture_output=True, text=True,
                         timeout=1200, env=dict(os.environ, PYTHONPATH=ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    R = np.load(npz)
    for dim, p, q, L, nsrc, seed in cases:
        pb, bdry = seeded_problem(dim, p, q, L, nsrc, seed)
        if dim == 20:
            ls, mg, dp = orc.local_solve_stage_uniform_2D_ItI, orc.merge_stage_uniform_2D_ItI, orc.down_pass_uniform_2D_ItI
        elif dim == 3:
            ls, mg, dp = orc.local_solve_stage_uniform_3D_DtN, orc.merge_stage_uniform_3D_DtN, orc.down_pass_uniform_3D_DtN
        else:
            ls, mg, dp = orc.local_solve_stage_uniform_2D_DtN, orc.merge_stage_uniform_2D_DtN, orc.down_pass_uniform_2D_DtN
        Y, T, v, h = ls(pb)
        S_lst, g_lst, T_top = mg(T, h, L, return_T=True)
        u = dp(bdry, S_lst, g_lst, Y, v)
        tag = f"{dim}_{p}_{q}_{L}_{nsrc}_{seed}"
        tol = 1e-9 if dim == 20 else 1e-11
        assert rel_err(h, R[tag + ":h"]) < tol and rel_err(T_top, R[tag + ":T_top"]) < tol, tag
        assert rel_err(g_lst[-1], R[tag + ":g_root"]) < tol and rel_err(u, R[tag + ":u"]) < tol, tag
