"""``solve``: mirror of `src/jaxhps/_solve.py:18-93` for uniform DtN problems."""
from __future__ import annotations

import numpy as np
import torch

from ._pdeproblem import PDEProblem
from .down_pass import down_pass_uniform_2D_DtN, down_pass_uniform_2D_ItI, down_pass_uniform_3D_DtN


def solve(pde_problem: PDEProblem, boundary_data, source=None, compute_device=None, host_device=None):
    """Downward pass with the operators stored by :func:`build_solver`.  Returns the solution on
    the HPS grid, shape ``(n_leaves, p^d[, n_src])``."""
    if source is not None:
        if not pde_problem.domain.bool_2D or not pde_problem.domain.bool_uniform:
            raise ValueError(
                "Source can only be specified for 2D uniform ItI problems. For other problems, the source must "
                "be specified at the time the solver is built."
            )
        return _up_then_down_pass(pde_problem, boundary_data, source, compute_device, host_device)
    if not pde_problem.domain.bool_uniform:
        # adaptive trees: boundary data is a list with one array per side / face (`_solve.py:96-112`)
        from .adaptive import down_pass_adaptive_2D_DtN, down_pass_adaptive_3D_DtN

        if not isinstance(boundary_data, list):
            raise ValueError(
                "For adaptive solves, boundary data needs to be a list. Try using the "
                "Domain.get_adaptive_boundary_data_lst() utility."
            )
        down = down_pass_adaptive_2D_DtN if pde_problem.domain.bool_2D else down_pass_adaptive_3D_DtN
        return down(pde_problem, boundary_data, device=compute_device, host_device=host_device)
    if isinstance(boundary_data, list):
        if all(isinstance(b, torch.Tensor) for b in boundary_data):
            boundary_data = torch.cat(boundary_data)
        else:
            boundary_data = np.concatenate([np.asarray(b) for b in boundary_data])
    if pde_problem.use_ItI:
        down = down_pass_uniform_2D_ItI
    else:
        down = down_pass_uniform_2D_DtN if pde_problem.domain.bool_2D else down_pass_uniform_3D_DtN
    return down(boundary_data, pde_problem.S_lst, pde_problem.g_tilde_lst, pde_problem.Y, pde_problem.v,
                device=compute_device, host_device=host_device)


def _up_then_down_pass(pde_problem: PDEProblem, boundary_data, source, compute_device, host_device):
    """Upward pass for the new source, then the ordinary downward pass (reference `_solve.py:115-151`)."""
    from . import _lib
    from .up_pass import up_pass_uniform_2D_DtN, up_pass_uniform_2D_ItI

    dev = _lib.require_cuda(compute_device)
    if pde_problem.use_ItI:
        up, down = up_pass_uniform_2D_ItI, down_pass_uniform_2D_ItI
    else:
        up, down = up_pass_uniform_2D_DtN, down_pass_uniform_2D_DtN
    v, g_tilde_lst = up(source, pde_problem, device=dev, host_device=dev)
    return down(boundary_data, pde_problem.S_lst, g_tilde_lst, pde_problem.Y, v, device=dev, host_device=host_device)
