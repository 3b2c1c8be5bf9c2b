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

# ---- reference-signature adapters for helpers whose shape differs here (test infrastructure only) ----
import numpy as _np  # noqa: E402


def _bounds_of(nodes, dim):
    keys = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[: 2 * dim]
    return _np.array([[float(getattr(n, k)) for k in keys] for n in nodes])


def _uniform_leaves(dim):
    def fn(root, L):
        b = (_grid.uniform_leaf_bounds_2D if dim == 2 else _grid.uniform_leaf_bounds_3D)(root, L)
        cls = _tree.DiscretizationNode2D if dim == 2 else _tree.DiscretizationNode3D
        return [cls(*row) for row in b]

    return fn


def _gauss_face(bounds, gauss_pts_1d):
    return _grid._gauss_panels_2D(_np.asarray(bounds, dtype=float)[None], len(gauss_pts_1d))


def _refinement_indexing(dim):
    def fn(p):
        r = (_grid.rearrange_indices_ext_int_2D if dim == 2 else _grid.rearrange_indices_ext_int_3D)(p)
        if dim == 2:
            ix, iy = _np.divmod(_np.arange(4 * p * p), 2 * p)
            blocks = [(ix < p) & (iy >= p), (ix >= p) & (iy >= p), (ix >= p) & (iy < p), (ix < p) & (iy < p)]
        else:
            ii = _np.arange(8 * p**3)
            ix, iy, iz = ii // (4 * p * p), (ii // (2 * p)) % (2 * p), ii % (2 * p)
            blocks = [((ix >= p) == xh) & ((iy >= p) == yh) & ((iz >= p) == zh) for zh in (True, False)
                      for xh, yh in ((False, False), (True, False), (True, True), (False, True))]
        return _np.concatenate([_np.flatnonzero(b)[r] for b in blocks]), r

    return fn


_alias("_grid_creation_2D", _grid, get_all_leaves=_tree.get_all_leaves,
       vmapped_bounds_2D=_grid.quad_children_bounds,
       bounds_for_quad_subdivision=lambda b: _grid.quad_children_bounds(_np.asarray(b, dtype=float)[None])[0],
       bounds_to_cheby_points_2D=lambda b, c: _grid.bounds_to_cheby_points_2D(_np.asarray(b, dtype=float)[None], len(c))[0],
       rearrange_indices_ext_int=_grid.rearrange_indices_ext_int_2D, get_all_uniform_leaves_2D=_uniform_leaves(2))
_alias("_grid_creation_3D", _grid, get_all_leaves=_tree.get_all_leaves,
       vmapped_bounds_3D=_grid.oct_children_bounds,
       bounds_for_oct_subdivision=lambda b: _grid.oct_children_bounds(_np.asarray(b, dtype=float)[None])[0],
       bounds_to_cheby_points_3D=lambda b, c: _grid.bounds_to_cheby_points_3D(_np.asarray(b, dtype=float)[None], len(c))[0],
       bounds_to_gauss_face=_gauss_face, rearrange_indices_ext_int=_grid.rearrange_indices_ext_int_3D,
       get_all_uniform_leaves_3D=_uniform_leaves(3))
_alias("_precompute_operators_3D", _operators, get_face_1_idxes=lambda p: _grid.face_cheby_indices_3D(p)[0],
       indexing_for_refinement_operator=_refinement_indexing(3))
_alias("_precompute_operators_2D", _operators, indexing_for_refinement_operator=_refinement_indexing(2))
_alias("_interpolation_methods", _interpolation_methods,
       interp_from_hps_2D=lambda leaves, p, f_evals, x_vals, y_vals: _interpolation_methods.interp_from_hps_2D(
           _bounds_of(leaves, 2), p, _np.asarray(f_evals), x_vals, y_vals),
       interp_from_hps_3D=lambda leaves, p, f_evals, x_vals, y_vals, z_vals: _interpolation_methods.interp_from_hps_3D(
           _bounds_of(leaves, 3), p, _np.asarray(f_evals), x_vals, y_vals, z_vals),
       interp_to_single_Chebyshev_panel_2D=lambda node_bounds, samples, p, from_x, from_y: _interpolation_methods.interp_to_hps_2D(
           _np.asarray(node_bounds, dtype=float)[None], _np.asarray(samples), p, from_x, from_y)[0],
       interp_to_single_Chebyshev_panel_3D=lambda node_bounds, samples, p, from_x, from_y, from_z:
       _interpolation_methods.interp_to_hps_3D(_np.asarray(node_bounds, dtype=float)[None], _np.asarray(samples), p, from_x,
                                               from_y, from_z)[0])
