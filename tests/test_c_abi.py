"""The C-ABI library loads and exports every symbol include/hps_b200.h declares; the product
path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import jaxhps_b200 as hps
from jaxhps_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.library_path())
    names = _declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in hps_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_version_and_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.hps_version() >= 100
    need = ctypes.c_size_t()
    assert lib.hps_local_solve_dtn_workspace(3, 10, 12, 10, 1, ctypes.byref(need)) == 0
    assert need.value > 10 * (1000 * 1000 + 1000 * 728) * 8
    assert lib.hps_merge_oct_dtn_level_workspace(8, 100, 1, ctypes.byref(need)) == 0
    assert need.value > 8 * 1200 * 1200 * 8
    assert lib.hps_local_solve_dtn_workspace(4, 10, 12, 10, 1, ctypes.byref(need)) < 0
    assert b"dim" in lib.hps_last_error_string()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    dom = hps.Domain(4, 2, root, 1)
    s = np.zeros((8, 64))
    pb = hps.PDEProblem(dom, source=s, D_xx_coefficients=s + 1, D_yy_coefficients=s + 1, D_zz_coefficients=s + 1)
    with pytest.raises(_lib.HpsLibraryError, match="no CPU fallback"):
        hps.build_solver(pb)
    # the adaptive, subtree-recomputation and sharded entry points fail the same way
    from jaxhps_b200._tree import add_eight_children

    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    add_eight_children(root, root=root, q=2)
    add_eight_children(root.children[0], root=root, q=2)
    dom = hps.Domain(4, 2, root)
    s = np.zeros((dom.n_leaves, 64))
    pb = hps.PDEProblem(dom, source=s, D_xx_coefficients=s + 1, D_yy_coefficients=s + 1, D_zz_coefficients=s + 1)
    with pytest.raises(_lib.HpsLibraryError, match="no CPU fallback"):
        hps.build_solver(pb)
    with pytest.raises(_lib.HpsLibraryError, match="no CPU fallback"):
        hps.solve(pb, dom.get_adaptive_boundary_data_lst(lambda x: x[..., 0]))
    from jaxhps_b200 import _dist_adaptive

    with pytest.raises(_lib.HpsLibraryError, match="no CPU fallback"):
        _dist_adaptive.CudaAdaptiveOps(None)
    dom2 = hps.Domain(4, 2, hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0), 2)
    s2 = np.zeros((16, 16))
    pb2 = hps.PDEProblem(dom2, source=s2, D_xx_coefficients=s2 + 1, D_yy_coefficients=s2 + 1)
    with pytest.raises(_lib.HpsLibraryError, match="no CPU fallback"):
        hps.solve_subtree(pb2, np.zeros(dom2.boundary_points.shape[0]), subtree_height=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jaxhps_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
