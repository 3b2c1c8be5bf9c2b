"""Top-level coupling of an ItI solver to an exterior boundary-integral formulation — SURVEY §8(f).1.

API mirror of the reference's flagship application, `examples/wave_scattering_utils.py`:
``get_DtN_from_ItI`` (:31-49), ``get_uin`` (:196-206), ``get_uin_and_normals`` (:130-193),
``setup_scattering_lin_system`` (:96-127) and ``get_scattering_uscat_impedance`` (:209-242).  The dense complex
algebra runs on the device through the C ABI — ``hps_zgesv`` (real embedding + pivoted FP64 LU) where the
reference calls ``jnp.linalg.solve`` and ``hps_zgemm_strided_batched`` (DMMA) where it writes ``@``; the plane-wave
data (n x n_src exponentials) is evaluated on the host.  The single- and double-layer matrices ``S``, ``D`` are inputs
(the reference loads them from MATLAB files).  There is no CPU fallback."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

_C = torch.complex128


def _dev_c(x, dev):
    return _lib.to_device(x, dev, dtype=_C)


def _zsolve(A: torch.Tensor, B: torch.Tensor, dev) -> torch.Tensor:
    """``A^-1 B`` on the device (``hps_zgesv``); A, B complex128, not modified."""
    lib = _lib.load()
    n = A.shape[0]
    B2 = B.reshape(n, -1).contiguous()
    A = A.contiguous()
    nrhs = B2.shape[1]
    X = torch.empty((n, nrhs), dtype=_C, device=dev)
    need = ctypes.c_size_t()
    _lib.check(lib.hps_zgesv_workspace(n, nrhs, ctypes.byref(need)), "hps_zgesv_workspace")
    ws = _lib.workspace(need.value, dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.hps_zgesv(_lib.stream_ptr(), n, nrhs, A.data_ptr(), n, B2.data_ptr(), nrhs, X.data_ptr(), ws.data_ptr(),
                             ws.numel(), info.data_ptr()), "hps_zgesv")
    _lib.check_info(info, "complex solve")
    return X.reshape(B.shape)


def _zmm(A: torch.Tensor, B: torch.Tensor, dev) -> torch.Tensor:
    """``A @ B`` on the device (``hps_zgemm_strided_batched``)."""
    lib = _lib.load()
    A = A.contiguous()
    M, K = A.shape
    B2 = B.reshape(K, -1).contiguous()
    N = B2.shape[1]
    Cm = torch.empty((M, N), dtype=_C, device=dev)
    ws = _lib.workspace(4 * K * N * 8, dev)
    _lib.check(lib.hps_zgemm_strided_batched(_lib.stream_ptr(), M, N, K, 1.0, A.data_ptr(), K, 0, B2.data_ptr(), 0, 0.0,
                                             Cm.data_ptr(), N, 0, 1, ws.data_ptr()), "hps_zgemm_strided_batched")
    return Cm.reshape((M,) + tuple(B.shape[1:]))


def get_DtN_from_ItI(R, eta: float, device=None, host_device=None):
    """``T = -i eta (R - I)^-1 (R + I)`` (eq. 2.17 of Gillman, Barnett, Martinsson; reference
    `wave_scattering_utils.py:31-49`).  ``R`` (n, n) complex128."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Rd = _dev_c(R, dev)
        eye = torch.eye(Rd.shape[0], dtype=_C, device=dev)
        T = _zsolve(Rd - eye, Rd + eye, dev)
        T.mul_(-1j * eta)
        return _lib.to_result(T, host_device)


def get_uin(k: float, pts, source_directions):
    """Incoming plane waves ``exp(i k <x, s>)``, shape (n, n_sources) (`wave_scattering_utils.py:196-206`)."""
    pts = np.asarray(pts)
    th = np.atleast_1d(np.asarray(source_directions, dtype=float))
    vecs = np.stack([np.cos(th), np.sin(th)], axis=1)
    return np.exp(1j * k * (pts @ vecs.T))


def get_uin_and_normals(k: float, bdry_pts, source_directions):
    """``(uin, d uin / dn)`` on the boundary points, sides in the order S, E, N, W
    (`wave_scattering_utils.py:130-193`)."""
    bdry_pts = np.asarray(bdry_pts)
    nps = bdry_pts.shape[0] // 4
    uin = get_uin(k, bdry_pts, source_directions)
    th = np.atleast_1d(np.asarray(source_directions, dtype=float))
    sx, sy = np.cos(th)[None, :], np.sin(th)[None, :]
    normals = np.concatenate([-1j * k * sy * uin[:nps], 1j * k * sx * uin[nps:2 * nps],
                              1j * k * sy * uin[2 * nps:3 * nps], -1j * k * sx * uin[3 * nps:]])
    return uin, normals


def setup_scattering_lin_system(S, D, T_int, gauss_bdry_pts, k: float, source_directions, device=None, host_device=None):
    """BIE system (3.4): ``A = I/2 - D + S T_int``, ``b = S (du_in/dn - T_int u_in)``
    (`wave_scattering_utils.py:96-127`)."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Sd, Dd, Td = _dev_c(S, dev), _dev_c(D, dev), _dev_c(T_int, dev)
        uin, normals = get_uin_and_normals(k, gauss_bdry_pts, source_directions)
        uin_d, nrm_d = _dev_c(uin, dev), _dev_c(normals, dev)
        n = Sd.shape[0]
        A = _zmm(Sd, Td, dev)
        A -= Dd
        A += 0.5 * torch.eye(n, dtype=_C, device=dev)
        b = _zmm(Sd, nrm_d - _zmm(Td, uin_d, dev), dev)
        return _lib.to_result(A, host_device), _lib.to_result(b, host_device)


def get_scattering_uscat_impedance(S, D, T, source_dirs, bdry_pts, k: float, eta: float, device=None, host_device=None):
    """Incoming impedance data of the scattered field on the boundary: solve the BIE system for ``u_scat``, then
    ``du_scat/dn = T (u_scat + u_in) - du_in/dn`` (eq. 1.12) and ``imp = du_scat/dn + i eta u_scat``
    (`wave_scattering_utils.py:209-242`)."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Td = _dev_c(T, dev)
        A, b = setup_scattering_lin_system(S, D, Td, bdry_pts, k, source_dirs, device=dev, host_device=dev)
        uin, uin_dn = get_uin_and_normals(k, bdry_pts, source_dirs)
        uin_d, dn_d = _dev_c(uin, dev), _dev_c(uin_dn, dev)
        uscat = _zsolve(A, b, dev)
        uscat_dn = _zmm(Td, uscat + uin_d, dev) - dn_d
        return _lib.to_result(uscat_dn + 1j * eta * uscat, host_device)


# ----------------------------------------------------------------------------------------------------------------
# Tangents of the two coupling steps: together with ``adjoint.top_T_jvp`` and ``adjoint.solve_jvp`` they give the
# directional derivative of the reference's whole inverse-scattering forward model (coefficients -> root ItI operator
# -> DtN -> incoming impedance data -> solution; `examples/inverse_scattering_utils.py:110-171`), which the reference
# obtains from ``jax.jvp``.  Same device algebra as above (``hps_zgesv`` / ``hps_zgemm_strided_batched``).


def get_DtN_from_ItI_jvp(R, dR, eta: float, device=None, host_device=None):
    """``(T, dT)`` for ``T = -i eta (R - I)^-1 (R + I)``:  ``dT = -(R - I)^-1 dR (T + i eta I)``."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Rd, dRd = _dev_c(R, dev), _dev_c(dR, dev)
        n = Rd.shape[0]
        eye = torch.eye(n, dtype=_C, device=dev)
        T = _zsolve(Rd - eye, Rd + eye, dev)
        T.mul_(-1j * eta)
        dT = _zsolve(Rd - eye, _zmm(dRd, T + 1j * eta * eye, dev), dev)
        dT.neg_()
        return _lib.to_result(T, host_device), _lib.to_result(dT, host_device)


def get_scattering_uscat_impedance_jvp(S, D, T, dT, source_dirs, bdry_pts, k: float, eta: float, device=None, host_device=None):
    """``(imp, d imp)`` of :func:`get_scattering_uscat_impedance` for a tangent ``dT`` of the interior DtN map:
    with ``A = I/2 - D + S T``, ``u = A^-1 S (dn u_in - T u_in)``:
    ``du = -A^-1 S dT (u + u_in)``, ``d(du/dn) = dT (u + u_in) + T du``, ``d imp = d(du/dn) + i eta du``."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Sd, Td, dTd = _dev_c(S, dev), _dev_c(T, dev), _dev_c(dT, dev)
        A, b = setup_scattering_lin_system(Sd, D, Td, bdry_pts, k, source_dirs, device=dev, host_device=dev)
        uin, uin_dn = get_uin_and_normals(k, bdry_pts, source_dirs)
        uin_d, dn_d = _dev_c(uin, dev), _dev_c(uin_dn, dev)
        uscat = _zsolve(A, b, dev)
        tot = uscat + uin_d
        imp = _zmm(Td, tot, dev) - dn_d + 1j * eta * uscat
        dT_tot = _zmm(dTd, tot, dev)
        du = _zsolve(A, _zmm(Sd, dT_tot, dev), dev)
        du.neg_()
        dimp = dT_tot + _zmm(Td, du, dev) + 1j * eta * du
        return _lib.to_result(imp, host_device), _lib.to_result(dimp, host_device)


def get_DtN_from_ItI_vjp(R, T_bar, eta: float, device=None, host_device=None):
    """Cotangent of ``R`` for a cotangent ``T_bar`` of ``T = get_DtN_from_ItI(R)`` (plain transposes, no conjugation):
    ``dT = -(R - I)^-1 dR (T + i eta I)``  =>  ``R_bar = -(R - I)^-T T_bar (T + i eta I)^T``."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Rd, Tb = _dev_c(R, dev), _dev_c(T_bar, dev)
        n = Rd.shape[0]
        eye = torch.eye(n, dtype=_C, device=dev)
        T = _zsolve(Rd - eye, Rd + eye, dev)
        T.mul_(-1j * eta)
        X = _zmm(Tb, (T + 1j * eta * eye).T.contiguous(), dev)
        R_bar = _zsolve((Rd - eye).T.contiguous(), X, dev)
        R_bar.neg_()
        return _lib.to_result(R_bar, host_device)


def get_scattering_uscat_impedance_vjp(S, D, T, imp_bar, source_dirs, bdry_pts, k: float, eta: float, device=None,
                                       host_device=None):
    """Cotangent of the interior DtN map ``T`` for a cotangent ``imp_bar`` (n, n_src) of
    :func:`get_scattering_uscat_impedance`: with ``b_bar = A^-T (T^T imp_bar + i eta imp_bar)``,
    ``T_bar = (imp_bar - S^T b_bar) (u + u_in)^T``."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        Sd, Td, ib = _dev_c(S, dev), _dev_c(T, dev), _dev_c(imp_bar, dev)
        A, b = setup_scattering_lin_system(Sd, D, Td, bdry_pts, k, source_dirs, device=dev, host_device=dev)
        uin, _ = get_uin_and_normals(k, bdry_pts, source_dirs)
        uin_d = _dev_c(uin, dev)
        uscat = _zsolve(A, b, dev)
        tot = uscat + uin_d
        us_bar = _zmm(Td.T.contiguous(), ib, dev) + 1j * eta * ib
        b_bar = _zsolve(A.T.contiguous(), us_bar, dev)
        left = ib - _zmm(Sd.T.contiguous(), b_bar, dev)
        T_bar = _zmm(left, tot.T.contiguous(), dev)
        return _lib.to_result(T_bar, host_device)
