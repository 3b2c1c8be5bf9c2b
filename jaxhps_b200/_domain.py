"""``Domain``: discretisation parameters plus the leaf / boundary point clouds.

API mirror of `src/jaxhps/_domain.py:38-97` for uniform trees (``L`` given).  Adaptive
construction (``L=None`` / ``from_adaptive_discretization``) belongs to SURVEY §8(f) and is
not built yet; it raises ``NotImplementedError`` instead of silently mis-behaving.
"""
from __future__ import annotations

import numpy as np

from ._grid import (
    compute_boundary_Gauss_points_uniform_2D,
    compute_boundary_Gauss_points_uniform_3D,
    compute_interior_Chebyshev_points_uniform_2D,
    compute_interior_Chebyshev_points_uniform_3D,
)
from ._grid import uniform_leaf_bounds_2D, uniform_leaf_bounds_3D
from ._interpolation_methods import interp_from_hps_2D, interp_from_hps_3D, interp_to_hps_2D, interp_to_hps_3D
from ._tree import DiscretizationNode2D, DiscretizationNode3D


class Domain:
    def __init__(self, p: int, q: int, root, L: int | None = None):
        self.p = int(p)  #: Chebyshev points per dimension on a leaf
        self.q = int(q)  #: Gauss points per dimension on a leaf side/face
        self.root = root
        self.L = L
        self.bool_2D = isinstance(root, DiscretizationNode2D)
        if not self.bool_2D and not isinstance(root, DiscretizationNode3D):
            raise TypeError("root must be a DiscretizationNode2D or DiscretizationNode3D")
        if L is None:
            raise NotImplementedError(
                "adaptive discretisations (L=None) are outside the hot path built so far "
                "(SURVEY §8(f) item 2); pass the number of uniform refinement levels L"
            )
        self.bool_uniform = True
        if self.bool_2D:
            #: (n_leaves, p^2, 2)
            self.interior_points: np.ndarray = compute_interior_Chebyshev_points_uniform_2D(root, L, p)
            #: (4 * 2^L * q, 2)
            self.boundary_points: np.ndarray = compute_boundary_Gauss_points_uniform_2D(root, L, q)
            self.n_leaves = 4**L
        else:
            #: (n_leaves, p^3, 3)
            self.interior_points = compute_interior_Chebyshev_points_uniform_3D(root, L, p)
            #: (6 * 4^L * q^2, 3)
            self.boundary_points = compute_boundary_Gauss_points_uniform_3D(root, L, q)
            self.n_leaves = 8**L

    def _leaf_bounds(self) -> np.ndarray:
        fn = uniform_leaf_bounds_2D if self.bool_2D else uniform_leaf_bounds_3D
        return fn(self.root, self.L)

    def interp_to_interior_points(self, values, sample_points_x, sample_points_y, sample_points_z=None) -> np.ndarray:
        """Values on a regular grid (``meshgrid(..., indexing="ij")``) -> samples on the HPS grid,
        shape ``(n_leaves, p^d)`` (reference `_domain.py:99-208`)."""
        values = np.asarray(values)
        if values.ndim == 2:
            assert sample_points_z is None and values.shape == (len(sample_points_x), len(sample_points_y))
            return interp_to_hps_2D(self._leaf_bounds(), values, self.p, sample_points_x, sample_points_y)
        assert sample_points_z is not None
        assert values.shape == (len(sample_points_x), len(sample_points_y), len(sample_points_z))
        return interp_to_hps_3D(self._leaf_bounds(), values, self.p, sample_points_x, sample_points_y, sample_points_z)

    def interp_to_boundary_points(self, *args, **kwargs):
        raise NotImplementedError("interp_to_boundary_points is not implemented yet.")  # as in the reference

    def interp_from_interior_points(self, samples, eval_points_x, eval_points_y, eval_points_z=None):
        """Samples on the HPS grid ``(n_leaves, p^d)`` -> values on a regular grid and the target points
        (reference `_domain.py:218-294`)."""
        samples = np.asarray(samples)
        if self.bool_2D:
            assert eval_points_z is None
            return interp_from_hps_2D(self._leaf_bounds(), self.p, samples, eval_points_x, eval_points_y)
        assert eval_points_z is not None
        return interp_from_hps_3D(self._leaf_bounds(), self.p, samples, eval_points_x, eval_points_y, eval_points_z)

    @classmethod
    def from_adaptive_discretization(cls, *args, **kwargs):
        raise NotImplementedError(
            "adaptive mesh generation is out of scope for the hot path (SURVEY §2, §8(f))"
        )
