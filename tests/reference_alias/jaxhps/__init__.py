"""TEST INFRASTRUCTURE: the reference's package name mapped onto ``jaxhps_b200`` so that the reference's
own test files (host layer and public API) can be run against this package unmodified
(tests/test_reference_suite_on_this_package.py).  Internal helper modules of the reference whose functions
have a different shape here (vmapped single-box helpers) are not aliased."""
import sys
import types

import jaxhps_b200 as _b
from jaxhps_b200 import *  # noqa: F401,F403
from jaxhps_b200 import (
    _adaptive_discretization,
    _build_solver,
    _domain,
    _grid,
    _interpolation_methods,
    _operators,
    _pdeproblem,
    _solve,
    _subtree_recomp,
    _tree,
    quadrature,
)

__version__ = "0.2"


def _alias(name, mod, **extra):
    if extra:
        new = types.ModuleType("jaxhps." + name)
        new.__dict__.update({k: v for k, v in mod.__dict__.items() if not k.startswith("__")})
        new.__dict__.update(extra)
        mod = new
    sys.modules["jaxhps." + name] = mod


_alias("_domain", _domain)
_alias("_pdeproblem", _pdeproblem)
_alias("_build_solver", _build_solver)
_alias("_solve", _solve)
_alias("_subtree_recomp", _subtree_recomp)
_alias("_discretization_tree", _tree)
_alias("_discretization_tree_operations_2D", _tree)
_alias("_discretization_tree_operations_3D", _tree)
_alias("_adaptive_discretization_2D", _adaptive_discretization)
_alias("_adaptive_discretization_3D", _adaptive_discretization)
_alias("_grid_creation_2D", _grid, get_all_leaves=_tree.get_all_leaves)
_alias("_grid_creation_3D", _grid, get_all_leaves=_tree.get_all_leaves)
_alias("_precompute_operators_2D", _operators)
_alias("_precompute_operators_3D", _operators)
_alias("quadrature", quadrature)
_alias("_interpolation_methods", _interpolation_methods)
