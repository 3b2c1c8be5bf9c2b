"""``PDEProblem``: coefficients, source, constant operators and the solver state.

API mirror of `src/jaxhps/_pdeproblem.py:25-505` (same constructor arguments, same
attribute names, same validation errors).  Differences that are deliberate and local:

* arrays are NumPy (host) or ``torch`` CUDA tensors (device-resident results);
* the p^d x p^d second-derivative operators ``D_xx .. D_yz`` are built lazily: the CUDA
  leaf kernel rebuilds operator rows from the scaled 1-D matrix ``D1`` and never reads
  them, only the oracle and user code do;
* ``D1`` and ``half_side_len`` are exposed for the kernels.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from ._domain import Domain
from ._operators import (
    precompute_diff_operators_2D,
    precompute_G_2D_ItI,
    precompute_N_matrix_2D,
    precompute_N_tilde_matrix_2D,
    precompute_P_2D_DtN,
    precompute_P_2D_ItI,
    precompute_P_3D_DtN,
    precompute_Q_2D_DtN,
    precompute_Q_3D_DtN,
    precompute_QH_2D_ItI,
    precompute_projection_ops_2D,
    precompute_projection_ops_3D,
    scaled_diff_matrix_1D,
)
from ._grid import rearrange_indices_ext_int_3D
from ._tree import DiscretizationNode3D, get_all_leaves

_COEFF_NAMES = (
    "D_xx_coefficients",
    "D_xy_coefficients",
    "D_xz_coefficients",
    "D_yy_coefficients",
    "D_yz_coefficients",
    "D_zz_coefficients",
    "D_x_coefficients",
    "D_y_coefficients",
    "D_z_coefficients",
    "I_coefficients",
)

_SECOND_3D = {"D_xx": ("D_x", "D_x"), "D_yy": ("D_y", "D_y"), "D_zz": ("D_z", "D_z"),
              "D_xy": ("D_x", "D_y"), "D_xz": ("D_x", "D_z"), "D_yz": ("D_y", "D_z")}


def check_input_shapes(source, use_ItI: bool, expected_shape: Tuple[int, ...], **coeffs) -> None:
    """Same rule as `_pdeproblem.py:469-505`: every given coefficient array must have the
    shape of ``domain.interior_points[..., 0]``."""
    for name in _COEFF_NAMES:
        arr = coeffs.get(name)
        if arr is not None and tuple(arr.shape) != tuple(expected_shape):
            raise ValueError(
                f"{name} has shape {tuple(arr.shape)} but should have shape {tuple(expected_shape)} "
                "to match the Domain's interior points."
            )


class PDEProblem:
    def __init__(
        self,
        domain: Domain,
        source=None,
        D_xx_coefficients=None,
        D_xy_coefficients=None,
        D_xz_coefficients=None,
        D_yy_coefficients=None,
        D_yz_coefficients=None,
        D_zz_coefficients=None,
        D_x_coefficients=None,
        D_y_coefficients=None,
        D_z_coefficients=None,
        I_coefficients=None,
        use_ItI: bool = False,
        eta: float | None = None,
    ):
        self.domain = domain
        coeffs = dict(
            D_xx_coefficients=D_xx_coefficients,
            D_xy_coefficients=D_xy_coefficients,
            D_xz_coefficients=D_xz_coefficients,
            D_yy_coefficients=D_yy_coefficients,
            D_yz_coefficients=D_yz_coefficients,
            D_zz_coefficients=D_zz_coefficients,
            D_x_coefficients=D_x_coefficients,
            D_y_coefficients=D_y_coefficients,
            D_z_coefficients=D_z_coefficients,
            I_coefficients=I_coefficients,
        )
        # --- validation, in the reference's order (`_pdeproblem.py:47-82`)
        if isinstance(domain.root, DiscretizationNode3D):
            bool_2D = False
            if use_ItI:
                raise NotImplementedError("ItI merges are not supported for 3D problems.")
        else:
            bool_2D = True
            for name in ("D_xz_coefficients", "D_yz_coefficients", "D_zz_coefficients", "D_z_coefficients"):
                if coeffs[name] is not None:
                    raise ValueError("z coefficients can not be set for 2D problems.")
        if use_ItI and eta is None:
            raise ValueError("eta must be specified when using ItI merges.")
        if use_ItI and not domain.bool_uniform:
            raise ValueError("ItI merges are only supported for uniform 2D problems.")
        if source is None and (not bool_2D or not domain.bool_uniform):
            raise ValueError("Source must be specified for non-uniform or 3D problems.")
        check_input_shapes(
            source=source,
            use_ItI=use_ItI,
            expected_shape=domain.interior_points[..., 0].shape,
            **coeffs,
        )
        for name, val in coeffs.items():
            setattr(self, name, val)
        self.source = source
        self.use_ItI = bool(use_ItI)
        self.eta = eta

        if domain.bool_uniform:
            # uniform trees: every leaf has the same side, scale once (`_pdeproblem.py:129-138`)
            self.half_side_len = (domain.root.xmax - domain.root.xmin) / (2 ** (domain.L + 1))
        else:
            # adaptive trees: unit-scaled operators + per-leaf side lengths (`_pdeproblem.py:139-145`)
            self.half_side_len = 1.0
            self.sidelens = np.array([leaf.xmax - leaf.xmin for leaf in get_all_leaves(domain.root)], dtype=np.float64)
        p, q = domain.p, domain.q
        #: scaled 1-D Chebyshev differentiation matrix (p, p); input of the CUDA leaf kernel
        self.D1 = scaled_diff_matrix_1D(p, self.half_side_len)
        self._lazy = {}
        if bool_2D:
            self.D_x, self.D_y, self.D_xx, self.D_yy, self.D_xy = precompute_diff_operators_2D(
                p, self.half_side_len
            )
            if not use_ItI:
                self.P = precompute_P_2D_DtN(p, q)
                self.Q = precompute_Q_2D_DtN(p, q, self.D_x, self.D_y)
            else:
                self.P = precompute_P_2D_ItI(p, q)
                self.G = precompute_G_2D_ItI(precompute_N_tilde_matrix_2D(self.D_x, self.D_y, p), eta)
                self.QH = precompute_QH_2D_ItI(precompute_N_matrix_2D(self.D_x, self.D_y, p), p, q, eta)
            if not domain.bool_uniform:
                self.L_2f1, self.L_1f2 = precompute_projection_ops_2D(q)
        else:
            r = rearrange_indices_ext_int_3D(p)
            eye = np.eye(p)
            ix = np.ix_(r, r)
            self.D_x = np.kron(self.D1, np.kron(eye, eye))[ix]
            self.D_y = np.kron(eye, np.kron(self.D1, eye))[ix]
            self.D_z = np.kron(eye, np.kron(eye, self.D1))[ix]
            self.P = precompute_P_3D_DtN(p, q)
            self.Q = precompute_Q_3D_DtN(p, q, self.D_x, self.D_y, self.D_z)
            if not domain.bool_uniform:
                self.L_4f1, self.L_1f4 = precompute_projection_ops_3D(q)

        self.reset()

    # second-derivative operators of 3D problems are only materialised on demand
    def __getattr__(self, name):
        if name in _SECOND_3D and "_lazy" in self.__dict__ and not self.domain.bool_2D:
            if name not in self._lazy:
                a, b = _SECOND_3D[name]
                self._lazy[name] = getattr(self, a) @ getattr(self, b)
            return self._lazy[name]
        raise AttributeError(name)

    def reset(self) -> None:
        """Drop the stored solution operators (`_pdeproblem.py:236-247`)."""
        self.Y = None
        self.v = None
        self.S_lst: List = []
        self.g_tilde_lst: List = []
        self.D_inv_lst: List = []
        self.BD_inv_lst: List = []
        self.Phi = None

    def update_coefficients(self, source=None, **coeffs) -> None:
        """Replace coefficients / source and reset the solver state
        (`_pdeproblem.py:249-333`)."""
        unknown = set(coeffs) - set(_COEFF_NAMES)
        if unknown:
            raise TypeError(f"unknown coefficient argument(s): {sorted(unknown)}")
        self.reset()
        check_input_shapes(
            source=self.source,
            use_ItI=self.use_ItI,
            expected_shape=self.domain.interior_points[..., 0].shape,
            **coeffs,
        )
        if source is not None:
            self.source = source
        for name, val in coeffs.items():
            if val is not None:
                setattr(self, name, val)


def _get_PDEProblem_chunk(pde_problem: PDEProblem, start_idx: int, end_idx: int) -> PDEProblem:
    """A shallow view of ``pde_problem`` whose per-leaf arrays are sliced to
    ``[start_idx:end_idx)``; the constant operators are shared
    (`_pdeproblem.py:336-466`)."""
    new = PDEProblem.__new__(PDEProblem)
    new.__dict__.update(pde_problem.__dict__)
    for name in _COEFF_NAMES + ("source",):
        val = getattr(pde_problem, name)
        setattr(new, name, None if val is None else val[start_idx:end_idx])
    if not pde_problem.domain.bool_uniform:
        new.sidelens = pde_problem.sidelens[start_idx:end_idx]
    return new
