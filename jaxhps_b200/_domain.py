"""``Domain``: discretisation parameters plus the leaf / boundary point clouds.

API mirror of `src/jaxhps/_domain.py:38-434`: uniform trees (``L`` given) and adaptive trees
(``L=None``; ``from_adaptive_discretization`` builds one from a function and a tolerance).
"""
from __future__ import annotations

import numpy as np

from ._adaptive_discretization import (
    generate_adaptive_mesh_level_restriction_2D,
    generate_adaptive_mesh_level_restriction_3D,
)
from ._grid import (
    compute_boundary_Gauss_points_adaptive_2D,
    compute_boundary_Gauss_points_adaptive_3D,
    compute_interior_Chebyshev_points_adaptive_2D,
    compute_interior_Chebyshev_points_adaptive_3D,
    leaf_bounds,
    compute_boundary_Gauss_points_uniform_2D,
    compute_boundary_Gauss_points_uniform_3D,
    compute_interior_Chebyshev_points_uniform_2D,
    compute_interior_Chebyshev_points_uniform_3D,
)
from ._grid import uniform_leaf_bounds_2D, uniform_leaf_bounds_3D
from ._interpolation_methods import interp_from_hps_2D, interp_from_hps_3D, interp_to_hps_2D, interp_to_hps_3D
from ._tree import DiscretizationNode2D, DiscretizationNode3D, get_all_leaves


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
            # adaptive tree: whatever leaves `root` currently has (`_domain.py:82-97`)
            self.bool_uniform = False
            self.n_leaves = len(get_all_leaves(root))
            if self.bool_2D:
                self.interior_points = compute_interior_Chebyshev_points_adaptive_2D(root, p)
                self.boundary_points = compute_boundary_Gauss_points_adaptive_2D(root, q)
            else:
                self.interior_points = compute_interior_Chebyshev_points_adaptive_3D(root, p)
                self.boundary_points = compute_boundary_Gauss_points_adaptive_3D(root, q)
            return
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
        if not self.bool_uniform:
            return leaf_bounds(self.root)
        fn = uniform_leaf_bounds_2D if self.bool_2D else uniform_leaf_bounds_3D
        return fn(self.root, self.L)

    def interp_to_interior_points(self, values, sample_points_x, sample_points_y, sample_points_z=None, device=None,
                                  host_device=None):
        """Values on a regular grid (``meshgrid(..., indexing="ij")``) -> samples on the HPS grid,
        shape ``(n_leaves, p^d)`` (reference `_domain.py:99-208`).  ``device=<cuda device>`` runs the CUDA kernels
        (``hps_interp_to_hps``) and returns a tensor on that device (NumPy when ``host_device="cpu"``)."""
        if device is not None:
            from ._interpolation_methods import interp_to_hps_device

            return interp_to_hps_device(self._leaf_bounds(), values, self.p, sample_points_x, sample_points_y,
                                        sample_points_z, device=device, host_device=host_device)
        values = np.asarray(values)
        if values.ndim == 2:
            assert sample_points_z is None and values.shape == (len(sample_points_x), len(sample_points_y))
            return interp_to_hps_2D(self._leaf_bounds(), values, self.p, sample_points_x, sample_points_y)
        assert sample_points_z is not None
        assert values.shape == (len(sample_points_x), len(sample_points_y), len(sample_points_z))
        return interp_to_hps_3D(self._leaf_bounds(), values, self.p, sample_points_x, sample_points_y, sample_points_z)

    def interp_to_boundary_points(self, *args, **kwargs):
        raise NotImplementedError("interp_to_boundary_points is not implemented yet.")  # as in the reference

    def interp_from_interior_points(self, samples, eval_points_x, eval_points_y, eval_points_z=None, device=None,
                                    host_device=None):
        """Samples on the HPS grid ``(n_leaves, p^d)`` -> values on a regular grid and the target points
        (reference `_domain.py:218-294`).  ``device=<cuda device>`` runs the CUDA kernel (``hps_interp_from_hps``);
        the values then stay on that device unless ``host_device="cpu"``."""
        if device is not None:
            from ._interpolation_methods import interp_from_hps_device

            assert (eval_points_z is None) == self.bool_2D
            return interp_from_hps_device(self._leaf_bounds(), self.p, samples, eval_points_x, eval_points_y, eval_points_z,
                                          device=device, host_device=host_device)
        samples = np.asarray(samples)
        if self.bool_2D:
            assert eval_points_z is None
            return interp_from_hps_2D(self._leaf_bounds(), self.p, samples, eval_points_x, eval_points_y)
        assert eval_points_z is not None
        return interp_from_hps_3D(self._leaf_bounds(), self.p, samples, eval_points_x, eval_points_y, eval_points_z)

    def get_adaptive_boundary_data_lst(self, f) -> list:
        """Evaluate ``f`` ([..., d] -> [...]) on the boundary points, one array per side (2D: S,E,N,W)
        or face (3D: x-,x+,y-,y+,z-,z+) — the input format of the adaptive down pass
        (`_domain.py:296-365`)."""
        b, r = self.boundary_points, self.root
        if self.bool_2D:
            masks = [b[:, 1] == r.ymin, b[:, 0] == r.xmax, b[:, 1] == r.ymax, b[:, 0] == r.xmin]
        else:
            masks = [b[:, 0] == r.xmin, b[:, 0] == r.xmax, b[:, 1] == r.ymin, b[:, 1] == r.ymax,
                     b[:, 2] == r.zmin, b[:, 2] == r.zmax]
        return [np.asarray(f(b[m])) for m in masks]

    @classmethod
    def from_adaptive_discretization(cls, p: int, q: int, root, f, tol: float, use_level_restriction: bool = True,
                                     use_l_2_norm: bool = False, device=None) -> "Domain":
        """Refine ``root`` until ``f`` (one callable or a list, [..., d] -> [...]) is resolved to ``tol``
        in the relative L_inf (default) or L_2 sense, then build the Domain (`_domain.py:367-434`).
        ``device`` (extension): a CUDA device evaluates the refinement criterion of every round there
        (``hps_refine_check``); ``None`` keeps the reference's host evaluation."""
        fns = f if isinstance(f, list) else [f]
        gen = (generate_adaptive_mesh_level_restriction_2D if isinstance(root, DiscretizationNode2D)
               else generate_adaptive_mesh_level_restriction_3D)
        for fn in fns:
            gen(root=root, f_fn=fn, tol=tol, p=p, q=q, restrict_bool=use_level_restriction, l2_norm=use_l_2_norm, device=device)
        return cls(p=p, q=q, root=root)
