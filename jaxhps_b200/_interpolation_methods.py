"""Interpolation between regular grids and the HPS grid: host side (NumPy) and device side (CUDA, ``*_device``).

Behavioural restatement of `src/jaxhps/_interpolation_methods.py:24-340` — the pre/post-processing
step on either side of the build+solve path in every example.  The tensor
structure of the barycentric matrices is used directly (two/three 1-D factors per leaf or per
target point) instead of forming the Kronecker products the reference builds.
"""
from __future__ import annotations


import numpy as np

from ._grid import rearrange_indices_ext_int_2D, rearrange_indices_ext_int_3D
from .quadrature import _bary_factor, chebyshev_points

_EPS = np.finfo(np.float64).eps


def _factor_rows(from_pts: np.ndarray, to_pts: np.ndarray) -> np.ndarray:
    """(n_to, n_from) 1-D barycentric factor with the multi-D convention of the reference (exact
    coincidences are nudged by machine epsilon, `quadrature/_interpolation.py:189-194`)."""
    dist, w, norm = _bary_factor(from_pts, to_pts, eps_guard=True)
    return 1.0 / (dist.T * w[None, :] * norm[:, None])


def _cheb_nodes(lo: np.ndarray, hi: np.ndarray, p: int) -> np.ndarray:
    c = chebyshev_points(p)
    return 0.5 * (hi - lo)[:, None] * c[None, :] + 0.5 * (lo + hi)[:, None]


def _batched_factor(nodes: np.ndarray, t: np.ndarray) -> np.ndarray:
    """nodes (n, p), t (n,) -> (n, p): barycentric row from each node set to its own target."""
    p = nodes.shape[1]
    diff = nodes[:, :, None] - nodes[:, None, :]
    diff[:, np.arange(p), np.arange(p)] = 1.0
    w = np.prod(diff, axis=1)  # same convention as quadrature._bary_weights_inv, per row
    dist = t[:, None] - nodes
    dist = np.where(dist == 0, _EPS, dist)
    inv = 1.0 / (w * dist)
    return inv / inv.sum(axis=1, keepdims=True)


def _owning_leaf(pts: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    """First leaf (in storage order) whose closed box contains each point
    (`_interpolation_methods.py:49-60`); points outside every leaf map to leaf 0 like argmax does."""
    inside = np.ones((pts.shape[0], lo.shape[0]), dtype=bool)
    for d in range(pts.shape[1]):
        inside &= (pts[:, d, None] >= lo[None, :, d]) & (pts[:, d, None] <= hi[None, :, d])
    return np.argmax(inside, axis=1)


def interp_from_hps_2D(leaf_bounds: np.ndarray, p: int, f_evals: np.ndarray, x_vals: np.ndarray, y_vals: np.ndarray):
    """Evaluate the piecewise polynomial given by ``f_evals (n_leaves, p^2[, n_src])`` on the grid
    ``x_vals x y_vals``.  Returns ``(vals, target_pts)`` with the reference's (quirky) conventions:
    the targets come from ``meshgrid(x, y)`` (xy indexing) and ``vals`` is that point list reshaped to
    ``(n_x, n_y)`` (`_interpolation_methods.py:24-93`)."""
    x_vals, y_vals = np.asarray(x_vals, float), np.asarray(y_vals, float)
    n_x, n_y = x_vals.shape[0], y_vals.shape[0]
    X, Y = np.meshgrid(x_vals, y_vals)
    target_pts = np.stack([X, Y], axis=2)
    pts = target_pts.reshape(-1, 2)
    b = np.asarray(leaf_bounds, float)
    lo, hi = b[:, [0, 2]], b[:, [1, 3]]
    idx = _owning_leaf(pts, lo, hi)
    fx = _batched_factor(_cheb_nodes(lo[idx, 0], hi[idx, 0], p), pts[:, 0])
    fy = _batched_factor(_cheb_nodes(lo[idx, 1], hi[idx, 1], p)[:, ::-1], pts[:, 1])  # y stored descending
    f = np.asarray(f_evals)
    inv = np.empty(p * p, dtype=np.int64)
    inv[rearrange_indices_ext_int_2D(p)] = np.arange(p * p)
    f_nat = f[:, inv]  # natural (x slow, y fast-descending) order
    multi = f.ndim == 3
    fn = f_nat.reshape((f.shape[0], p, p) + f.shape[2:])[idx]
    if multi:
        vals = np.einsum("ni,nj,nijs->ns", fx, fy, fn).reshape(n_x, n_y, f.shape[-1])
    else:
        vals = np.einsum("ni,nj,nij->n", fx, fy, fn).reshape(n_x, n_y)
    return vals, target_pts


def interp_from_hps_3D(leaf_bounds: np.ndarray, p: int, f_evals: np.ndarray, x_vals, y_vals, z_vals):
    """3D analogue (`_interpolation_methods.py:156-219`); ``vals`` has shape ``(n_x, n_y, n_z)`` and the
    targets come from ``meshgrid(x, y, z)`` (xy indexing, as in the reference)."""
    x_vals, y_vals, z_vals = (np.asarray(a, float) for a in (x_vals, y_vals, z_vals))
    X, Y, Z = np.meshgrid(x_vals, y_vals, z_vals)
    target_pts = np.stack([X, Y, Z], axis=3)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)
    b = np.asarray(leaf_bounds, float)
    lo, hi = b[:, [0, 2, 4]], b[:, [1, 3, 5]]
    f = np.asarray(f_evals)
    inv = np.empty(p**3, dtype=np.int64)
    inv[rearrange_indices_ext_int_3D(p)] = np.arange(p**3)
    f_nat = f[:, inv].reshape(f.shape[0], p, p, p)
    out = np.empty(pts.shape[0], dtype=f.dtype)
    for s in range(0, pts.shape[0], 1 << 16):  # bounded temporaries
        q = pts[s : s + (1 << 16)]
        idx = _owning_leaf(q, lo, hi)
        fx = _batched_factor(_cheb_nodes(lo[idx, 0], hi[idx, 0], p), q[:, 0])
        fy = _batched_factor(_cheb_nodes(lo[idx, 1], hi[idx, 1], p), q[:, 1])
        fz = _batched_factor(_cheb_nodes(lo[idx, 2], hi[idx, 2], p), q[:, 2])
        out[s : s + q.shape[0]] = np.einsum("ni,nj,nk,nijk->n", fx, fy, fz, f_nat[idx], optimize=True)
    return out.reshape(x_vals.shape[0], y_vals.shape[0], z_vals.shape[0]), target_pts


def interp_to_hps_2D(leaf_bounds: np.ndarray, values: np.ndarray, p: int, from_x, from_y) -> np.ndarray:
    """Samples on a regular grid ``values (n_x, n_y)`` -> ``(n_leaves, p^2)`` on the HPS grid
    (`_interpolation_methods.py:278-308`)."""
    b = np.asarray(leaf_bounds, float)
    r = rearrange_indices_ext_int_2D(p)
    to_x = _cheb_nodes(b[:, 0], b[:, 1], p)
    to_y = _cheb_nodes(b[:, 2], b[:, 3], p)[:, ::-1]
    out = np.empty((b.shape[0], p * p), dtype=np.result_type(values, float))
    for leaf in range(b.shape[0]):
        Ix, Iy = _factor_rows(np.asarray(from_x, float), to_x[leaf]), _factor_rows(np.asarray(from_y, float), to_y[leaf])
        out[leaf] = (Ix @ values @ Iy.T).reshape(-1)[r]
    return out


def interp_to_hps_3D(leaf_bounds: np.ndarray, values: np.ndarray, p: int, from_x, from_y, from_z) -> np.ndarray:
    """3D analogue (`_interpolation_methods.py:311-340`)."""
    b = np.asarray(leaf_bounds, float)
    r = rearrange_indices_ext_int_3D(p)
    to = [_cheb_nodes(b[:, 2 * d], b[:, 2 * d + 1], p) for d in range(3)]
    out = np.empty((b.shape[0], p**3), dtype=np.result_type(values, float))
    fx, fy, fz = (np.asarray(a, float) for a in (from_x, from_y, from_z))
    for leaf in range(b.shape[0]):
        Ix, Iy, Iz = _factor_rows(fx, to[0][leaf]), _factor_rows(fy, to[1][leaf]), _factor_rows(fz, to[2][leaf])
        out[leaf] = np.einsum("ia,jb,kc,abc->ijk", Ix, Iy, Iz, values, optimize=True).reshape(-1)[r]
    return out


# =====================================================================================
# Device path (``hps_interp_from_hps`` / ``hps_interp_to_hps`` in libhps_b200.so; csrc/interp.cu)
# =====================================================================================


def _bary_weights_inv_host(x: np.ndarray) -> np.ndarray:
    from .quadrature import _bary_weights_inv

    return np.ascontiguousarray(_bary_weights_inv(np.asarray(x, dtype=np.float64)))


def _index_tables(p: int, dim: int):
    """(nat2leaf, leaf2nat): position in the leaf's storage order of each natural index, and the inverse."""
    r = rearrange_indices_ext_int_2D(p) if dim == 2 else rearrange_indices_ext_int_3D(p)
    leaf2nat = np.asarray(r, dtype=np.int32)
    nat2leaf = np.empty_like(leaf2nat)
    nat2leaf[leaf2nat] = np.arange(leaf2nat.shape[0], dtype=np.int32)
    return nat2leaf, leaf2nat


def interp_from_hps_device(leaf_bounds, p: int, f_evals, x_vals, y_vals, z_vals=None, device=None, host_device=None):
    """Device version of :func:`interp_from_hps_2D` / :func:`interp_from_hps_3D`: same return convention
    ``(vals, target_pts)``; ``vals`` is a torch tensor on the device unless ``host_device`` asks for NumPy."""
    import torch

    from . import _lib

    dev = _lib.require_cuda(device)
    lib = _lib.load()
    dim = 2 if z_vals is None else 3
    x_vals, y_vals = np.asarray(x_vals, float), np.asarray(y_vals, float)
    if dim == 2:
        X, Y = np.meshgrid(x_vals, y_vals)
        target_pts = np.stack([X, Y], axis=2)
        pts = target_pts.reshape(-1, 2)
        shape = (x_vals.shape[0], y_vals.shape[0])
    else:
        z_vals = np.asarray(z_vals, float)
        X, Y, Z = np.meshgrid(x_vals, y_vals, z_vals)
        target_pts = np.stack([X, Y, Z], axis=3)
        pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=-1)
        shape = (x_vals.shape[0], y_vals.shape[0], z_vals.shape[0])
    with torch.cuda.device(dev):
        f = _lib.to_device(f_evals, dev)
        multi = f.ndim == 3
        f3 = f if multi else f.unsqueeze(-1)
        n_leaves, npts_leaf, n_src = f3.shape
        nat2leaf, _ = _index_tables(p, dim)
        b = _lib.to_device(np.ascontiguousarray(np.asarray(leaf_bounds, float)), dev)
        cheb = _lib.to_device(chebyshev_points(p), dev)
        tbl = torch.from_numpy(nat2leaf).to(dev)
        pd = _lib.to_device(np.ascontiguousarray(pts), dev)
        out = torch.empty((pts.shape[0], n_src), dtype=torch.float64, device=dev)
        rc = lib.hps_interp_from_hps(_lib.stream_ptr(), dim, n_leaves, p, n_src, pts.shape[0], b.data_ptr(), cheb.data_ptr(),
                                     tbl.data_ptr(), f3.contiguous().data_ptr(), pd.data_ptr(), out.data_ptr())
        _lib.check(rc, "hps_interp_from_hps")
        vals = out.reshape(shape + (n_src,)) if multi else out[:, 0].reshape(shape)
        return _lib.to_result(vals, host_device if host_device is not None else dev), target_pts


def interp_to_hps_device(leaf_bounds, values, p: int, from_x, from_y, from_z=None, device=None, host_device=None):
    """Device version of :func:`interp_to_hps_2D` / :func:`interp_to_hps_3D`: ``(n_leaves, p^d)``."""
    import ctypes

    import torch

    from . import _lib

    dev = _lib.require_cuda(device)
    lib = _lib.load()
    dim = 2 if from_z is None else 3
    with torch.cuda.device(dev):
        V = _lib.to_device(values, dev)
        b_h = np.ascontiguousarray(np.asarray(leaf_bounds, float))
        n_leaves = b_h.shape[0]
        fr = [np.asarray(a, float) for a in ((from_x, from_y) if dim == 2 else (from_x, from_y, from_z))]
        if tuple(V.shape) != tuple(a.shape[0] for a in fr):
            raise ValueError(f"values of shape {tuple(V.shape)} do not match the sample grids {[a.shape[0] for a in fr]}")
        b = _lib.to_device(b_h, dev)
        cheb = _lib.to_device(chebyshev_points(p), dev)
        _, leaf2nat = _index_tables(p, dim)
        tbl = torch.from_numpy(leaf2nat).to(dev)
        fd = [_lib.to_device(np.ascontiguousarray(a), dev) for a in fr]
        wd = [_lib.to_device(_bary_weights_inv_host(a), dev) for a in fr]
        n = [a.shape[0] for a in fr] + ([1] if dim == 2 else [])
        out = torch.empty((n_leaves, p**dim), dtype=torch.float64, device=dev)
        # leaves are processed in slices so that the per-leaf intermediates stay bounded
        one = ctypes.c_size_t()
        _lib.check(lib.hps_interp_to_hps_workspace(dim, 1, p, n[0], n[1], n[2], ctypes.byref(one)), "workspace query")
        step = int(max(1, min(n_leaves, (4 << 30) // max(1, one.value), _lib.MAX_BATCH)))
        need = ctypes.c_size_t()
        _lib.check(lib.hps_interp_to_hps_workspace(dim, step, p, n[0], n[1], n[2], ctypes.byref(need)), "workspace query")
        ws = _lib.workspace(need.value, dev)
        for s0 in range(0, n_leaves, step):
            s1 = min(n_leaves, s0 + step)
            rc = lib.hps_interp_to_hps(_lib.stream_ptr(), dim, s1 - s0, p, n[0], n[1], n[2], b[s0:s1].data_ptr(), cheb.data_ptr(),
                                       fd[0].data_ptr(), fd[1].data_ptr(), fd[2].data_ptr() if dim == 3 else None,
                                       wd[0].data_ptr(), wd[1].data_ptr(), wd[2].data_ptr() if dim == 3 else None,
                                       tbl.data_ptr(), V.contiguous().data_ptr(), out[s0:s1].data_ptr(), ws.data_ptr(), ws.numel())
            _lib.check(rc, "hps_interp_to_hps")
        return _lib.to_result(out, host_device if host_device is not None else dev)
