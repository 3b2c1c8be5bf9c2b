"""Subtree recomputation for uniform 2D problems: build and solve without ever holding more than one
subtree's operators (API mirror of `src/jaxhps/_subtree_recomp.py:20-541`).

Per chunk of ``4**subtree_height`` leaves the local solves and merges are run keeping only the
subtree root's ``(T, h)``; the subtree roots are merged on top; on the way down every subtree is
rebuilt and its own down pass is run.  (The parallel version of this split is `_dist.py`.)
Deviation: the reference's partial down pass inside ``solve_subtree`` omits the ``Y_arr``/``v_arr``
arguments and would raise; here they are passed as ``None`` as intended."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._device_config import local_solve_chunksize_2D
from ._pdeproblem import PDEProblem, _get_PDEProblem_chunk
from .down_pass import down_pass_uniform_2D_DtN, down_pass_uniform_2D_ItI
from .local_solve import local_solve_stage_uniform_2D_DtN, local_solve_stage_uniform_2D_ItI
from .merge import merge_stage_uniform_2D_DtN, merge_stage_uniform_2D_ItI


def _fns(pde_problem):
    if pde_problem.use_ItI:
        return local_solve_stage_uniform_2D_ItI, merge_stage_uniform_2D_ItI, down_pass_uniform_2D_ItI
    return local_solve_stage_uniform_2D_DtN, merge_stage_uniform_2D_DtN, down_pass_uniform_2D_DtN


def _check(pde_problem):
    if not pde_problem.domain.bool_2D:
        raise ValueError("Subtree recomputation is only supported for 2D problems.")
    if not pde_problem.domain.bool_uniform:
        raise ValueError("Subtree recomputation is only supported for uniform quadtrees.")


def _cat(xs):
    return torch.cat(xs, dim=0) if isinstance(xs[0], torch.Tensor) else np.concatenate(xs, axis=0)


def _local_solve_and_build(pde_problem, boundary_data, subtree_height, compute_device, host_device, return_top_T=False):
    """Upward pass when ``boundary_data is None`` (returns the top-level merge output), otherwise the
    downward pass over the subtrees (returns the solution) (`_subtree_recomp.py:232-392`)."""
    local_solve_fn, merge_fn, down_pass_fn = _fns(pde_problem)
    dev = _lib.require_cuda(compute_device)
    n_leaves = pde_problem.domain.n_leaves
    chunk = 4**subtree_height
    if n_leaves % chunk:
        raise ValueError("4**subtree_height must divide the number of leaves")
    n_chunks = n_leaves // chunk
    upward = boundary_data is None
    T_lst, h_lst, solns = [], [], []
    for i, start in enumerate(range(0, n_leaves, chunk)):
        sub = _get_PDEProblem_chunk(pde_problem, start, start + chunk)
        Y, T, v, h = local_solve_fn(sub, device=dev, host_device=dev)
        out = merge_fn(T, h, l=subtree_height, device=dev, host_device=dev, subtree_recomp=upward)
        if upward:
            T_lst.append(out[0])
            h_lst.append(out[1])
        else:
            S_lst, g_lst = out
            per = boundary_data.shape[0] // n_chunks
            solns.append(down_pass_fn(boundary_data[i * per : (i + 1) * per], S_lst, g_lst, Y, v, device=dev, host_device=dev))
    if upward:
        return merge_fn(_cat(T_lst), _cat(h_lst), l=pde_problem.domain.L - subtree_height, device=dev, host_device=dev,
                        subtree_recomp=False, return_T=return_top_T)
    return _lib.to_result(_cat(solns), host_device)


def upward_pass_subtree(pde_problem: PDEProblem, subtree_height: int = 7, compute_device=None, host_device=None):
    """Build by subtrees; stores the top levels' ``S_lst``/``g_tilde_lst`` on the problem and returns the
    top-level Poincaré–Steklov matrix (`_subtree_recomp.py:395-456`)."""
    _check(pde_problem)
    out = _local_solve_and_build(pde_problem, None, subtree_height, compute_device, host_device, return_top_T=True)
    pde_problem.S_lst, pde_problem.g_tilde_lst = out[0], out[1]
    return _lib.to_result(out[2], host_device)


def downward_pass_subtree(pde_problem: PDEProblem, boundary_data, subtree_height: int = 7, compute_device=None,
                          host_device=None):
    """Propagate boundary data through the stored top levels, then rebuild and solve every subtree
    (`_subtree_recomp.py:459-541`)."""
    _check(pde_problem)
    if isinstance(boundary_data, list):
        boundary_data = np.concatenate([np.asarray(b) for b in boundary_data])
    _, _, down_pass_fn = _fns(pde_problem)
    dev = _lib.require_cuda(compute_device)
    bdry = down_pass_fn(boundary_data, pde_problem.S_lst, pde_problem.g_tilde_lst, None, None, device=dev, host_device=dev)
    return _local_solve_and_build(pde_problem, bdry, subtree_height, compute_device, host_device)


def solve_subtree(pde_problem: PDEProblem, boundary_data, subtree_height: int = 7, compute_device=None, host_device=None):
    """Build + solve in one call; small problems (≤ the 2D chunk size) are done in one piece exactly
    like the reference's ``_all_together_*`` (`_subtree_recomp.py:20-229`)."""
    _check(pde_problem)
    if isinstance(boundary_data, list):
        boundary_data = np.concatenate([np.asarray(b) for b in boundary_data])
    local_solve_fn, merge_fn, down_pass_fn = _fns(pde_problem)
    dev = _lib.require_cuda(compute_device)
    dtype = np.complex128 if pde_problem.use_ItI else np.float64
    n_leaves = pde_problem.domain.n_leaves
    if local_solve_chunksize_2D(pde_problem.domain.p, dtype) >= n_leaves and 4**subtree_height >= n_leaves:
        Y, T, v, h = local_solve_fn(pde_problem, device=dev, host_device=dev)
        S_lst, g_lst = merge_fn(T, h, l=pde_problem.domain.L, device=dev, host_device=dev)
        return down_pass_fn(boundary_data, S_lst, g_lst, Y, v, device=dev, host_device=host_device)
    S_lst, g_lst = _local_solve_and_build(pde_problem, None, subtree_height, compute_device, host_device)
    bdry = down_pass_fn(boundary_data, S_lst, g_lst, None, None, device=dev, host_device=dev)
    return _local_solve_and_build(pde_problem, bdry, subtree_height, compute_device, host_device)
