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
