"""``build_solver``: dispatch + result stashing, mirroring `src/jaxhps/_build_solver.py:38-171`.

The reference loops over leaf chunks on the host because its assembled operators do not
fit on an 80 GB device; here the chunking lives inside the local-solve stage (bounded scratch),
so the driver is a straight line."""
from __future__ import annotations

from ._pdeproblem import PDEProblem
from .local_solve import (
    local_solve_stage_uniform_2D_DtN,
    local_solve_stage_uniform_2D_ItI,
    local_solve_stage_uniform_3D_DtN,
)
from .merge import merge_stage_uniform_2D_DtN, merge_stage_uniform_2D_ItI, merge_stage_uniform_3D_DtN


def build_solver(pde_problem: PDEProblem, return_top_T: bool = False, compute_device=None, host_device=None):
    """Run the local solves and all merges; store ``Y, v, S_lst, g_tilde_lst`` on the problem.

    ``compute_device``: CUDA device (default: current).  ``host_device``: where the stored
    operators live — ``None``/"cpu" copies them to host NumPy arrays like the reference's
    default; a CUDA device keeps them resident (recommended: ``solve`` then moves nothing).
    Returns the top-level Poincaré–Steklov matrix when ``return_top_T`` is set."""
    if pde_problem.source is None:
        if not pde_problem.domain.bool_uniform or not pde_problem.domain.bool_2D:
            raise ValueError(
                "Build stage for problems without source terms is only implemented for 2D uniform ItI problems."
            )
        return _nosource_build_solver(pde_problem, return_top_T, compute_device, host_device)
    if not pde_problem.domain.bool_uniform:
        return _adaptive_build_solver(pde_problem, return_top_T, compute_device, host_device)
    from . import _lib

    # leaf outputs stay on the compute device between the two stages (the reference round-trips
    # them through the host, `_build_solver.py:136-169`, which its docs name as the bottleneck)
    dev = _lib.require_cuda(compute_device)
    if pde_problem.use_ItI:
        Y, T, v, h = local_solve_stage_uniform_2D_ItI(pde_problem, device=dev, host_device=dev)
        merge_fn = merge_stage_uniform_2D_ItI
    elif pde_problem.domain.bool_2D:
        Y, T, v, h = local_solve_stage_uniform_2D_DtN(pde_problem, device=dev, host_device=dev)
        merge_fn = merge_stage_uniform_2D_DtN
    else:
        Y, T, v, h = local_solve_stage_uniform_3D_DtN(pde_problem, device=dev, host_device=dev)
        merge_fn = merge_stage_uniform_3D_DtN
    pde_problem.Y = _lib.to_result(Y, host_device)
    pde_problem.v = _lib.to_result(v, host_device)
    out = merge_fn(T, h, l=pde_problem.domain.L, device=dev, host_device=host_device, return_T=return_top_T)
    pde_problem.S_lst = out[0]
    pde_problem.g_tilde_lst = out[1]
    if return_top_T:
        return out[2]
    return None


def _nosource_build_solver(pde_problem: PDEProblem, return_top_T: bool, compute_device, host_device):
    """Source-free build for 2D uniform problems: keeps ``Phi``, ``D_inv_lst`` and ``BD_inv_lst`` so
    that ``solve(..., source=f)`` can run an upward pass (reference `_build_solver.py:261-331`)."""
    from . import _lib
    from .up_pass import (
        nosource_local_solve_stage_uniform_2D_DtN,
        nosource_local_solve_stage_uniform_2D_ItI,
        nosource_merge_stage_uniform_2D_DtN,
        nosource_merge_stage_uniform_2D_ItI,
    )

    dev = _lib.require_cuda(compute_device)
    if pde_problem.use_ItI:
        ls, mg = nosource_local_solve_stage_uniform_2D_ItI, nosource_merge_stage_uniform_2D_ItI
    else:
        ls, mg = nosource_local_solve_stage_uniform_2D_DtN, nosource_merge_stage_uniform_2D_DtN
    Y, T, Phi = ls(pde_problem, device=dev, host_device=dev)
    pde_problem.Y = _lib.to_result(Y, host_device)
    pde_problem.Phi = _lib.to_result(Phi, host_device)
    out = mg(T, pde_problem.domain.L, device=dev, host_device=host_device, return_T=return_top_T)
    pde_problem.S_lst, pde_problem.D_inv_lst, pde_problem.BD_inv_lst = out[0], out[1], out[2]
    return out[3] if return_top_T else None


def _adaptive_build_solver(pde_problem: PDEProblem, return_top_T: bool, compute_device, host_device):
    """Adaptive trees (reference `_build_solver.py:174-258`): leaf solves, then node-by-node merges.
    Leaf outputs are also attached to ``leaf.data`` like the reference does; device copies of ``Y``
    and ``v`` stay on the problem so that ``solve`` moves nothing."""
    from . import _lib
    from ._tree import get_all_leaves
    from .adaptive import (
        local_solve_stage_adaptive_2D_DtN,
        local_solve_stage_adaptive_3D_DtN,
        merge_stage_adaptive_2D_DtN,
        merge_stage_adaptive_3D_DtN,
    )

    dev = _lib.require_cuda(compute_device)
    two_d = pde_problem.domain.bool_2D
    ls = local_solve_stage_adaptive_2D_DtN if two_d else local_solve_stage_adaptive_3D_DtN
    mg = merge_stage_adaptive_2D_DtN if two_d else merge_stage_adaptive_3D_DtN
    Y, T, v, h = ls(pde_problem, device=dev, host_device=dev)
    pde_problem.__dict__["_adaptive_leaf"] = (Y, v)
    pde_problem.Y = _lib.to_result(Y, host_device)
    pde_problem.v = _lib.to_result(v, host_device)
    T_res, h_res = _lib.to_result(T, host_device), _lib.to_result(h, host_device)
    mg(pde_problem, T, h, device=dev, host_device=host_device, return_T=return_top_T)
    for i, leaf in enumerate(get_all_leaves(pde_problem.domain.root)):
        leaf.data.Y, leaf.data.v = pde_problem.Y[i], pde_problem.v[i]
        leaf.data.T, leaf.data.h = T_res[i], h_res[i]
    if return_top_T:
        return pde_problem.domain.root.data.T
    return None
