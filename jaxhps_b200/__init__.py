"""jaxhps_b200 — the HPS build-and-solve hot path of meliao/jaxhps as sm_100a CUDA kernels behind the same
Python surface (reference `src/jaxhps/__init__.py:1-35`).

What lives where:

=====================  ==========================================================================
host layer (NumPy)     ``_tree`` (boxes, splitting), ``_domain`` (``Domain``), ``_pdeproblem``
                       (``PDEProblem``), ``_grid`` / ``_operators`` / ``quadrature`` (point clouds and
                       pre-computed operators), ``_adaptive_discretization`` (mesh generation),
                       ``_adaptive_plan`` (index tables of non-uniform trees)
stage functions        ``local_solve``, ``merge``, ``down_pass``, ``up_pass``, ``adaptive`` — the reference's
                       names and signatures; bodies are calls into ``libhps_b200.so`` (``_lib``)
drivers                ``build_solver`` / ``solve`` (``_build_solver``, ``_solve``), subtree recomputation
                       (``_subtree_recomp``), multi-GPU sharding (``_dist``, ``_dist_adaptive``)
derivatives            ``adjoint`` — ``solve_jvp`` / ``solve_vjp`` of the 2D uniform solve (what the reference gets from
                       ``jax.jvp`` / ``jax.vjp``), ``scattering`` — ItI -> DtN conversion and the BIE coupling solve
=====================  ==========================================================================

There is no CPU fallback: every stage raises ``_lib.HpsLibraryError`` without CUDA or without the library.
"""
from . import down_pass, local_solve, merge, quadrature, up_pass  # noqa: F401  (sub-modules named like the reference's packages)
from ._build_solver import build_solver
from ._device_config import local_solve_chunksize_2D, local_solve_chunksize_3D
from ._domain import Domain
from ._pdeproblem import PDEProblem
from ._solve import solve
from ._subtree_recomp import downward_pass_subtree, solve_subtree, upward_pass_subtree
from ._tree import DiscretizationNode2D, DiscretizationNode3D, get_all_leaves

__version__ = "0.1"

#: the reference's public names (`src/jaxhps/__init__.py:21-34`)
__all__ = sorted(
    n
    for n, obj in list(globals().items())
    if not n.startswith("_") and n not in ("down_pass", "local_solve", "merge", "quadrature", "up_pass") and callable(obj)
)
