"""NumPy restatement of the reference's ItI -> DtN conversion and BIE coupling (CPU oracle; TEST INFRASTRUCTURE, never
imported by the jaxhps_b200 package).  Follows /root/reference/examples/wave_scattering_utils.py:
``get_DtN_from_ItI`` :31-49, ``get_uin`` :196-206, ``get_uin_and_normals`` :130-193, ``setup_scattering_lin_system``
:96-127, ``get_scattering_uscat_impedance`` :209-242.  Third-party arithmetic: ``jnp.linalg.solve`` (jax/jaxlib,
unpinned, absent) -> ``numpy.linalg.solve`` (LAPACK zgesv).  Pinned to the reference's own functions executed on the
NumPy ``jax`` shim: tests/golden/make_golden_scattering.py -> tests/golden/scattering_reference.npz."""
import numpy as np


def get_DtN_from_ItI(R, eta):
    n = R.shape[0]
    eye = np.eye(n)
    return -1j * eta * np.linalg.solve(R - eye, R + eye)


def get_uin(k, pts, source_directions):
    vecs = np.array([np.cos(source_directions), np.sin(source_directions)]).T
    return np.exp(1j * k * np.dot(pts, vecs.T))


def get_uin_and_normals(k, bdry_pts, source_directions):
    nps = bdry_pts.shape[0] // 4
    uin = get_uin(k, bdry_pts, source_directions)
    vecs = np.array([np.cos(source_directions), np.sin(source_directions)]).T
    normals = np.concatenate([
        -1j * k * np.expand_dims(vecs[:, 1], axis=0) * uin[:nps],
        1j * k * np.expand_dims(vecs[:, 0], axis=0) * uin[nps:2 * nps],
        1j * k * np.expand_dims(vecs[:, 1], axis=0) * uin[2 * nps:3 * nps],
        -1j * k * np.expand_dims(vecs[:, 0], axis=0) * uin[3 * nps:],
    ])
    return uin, normals


def setup_scattering_lin_system(S, D, T_int, gauss_bdry_pts, k, source_directions):
    n = gauss_bdry_pts.shape[0]
    uin, normals = get_uin_and_normals(k, gauss_bdry_pts, source_directions)
    A = 0.5 * np.eye(n) - D + S @ T_int
    b = S @ (normals - T_int @ uin)
    return A, b


def get_scattering_uscat_impedance(S, D, T, source_dirs, bdry_pts, k, eta):
    A, b = setup_scattering_lin_system(S, D, T, bdry_pts, k, source_dirs)
    uin, uin_dn = get_uin_and_normals(k, bdry_pts, source_dirs)
    uscat = np.linalg.solve(A, b)
    uscat_dn = T @ (uscat + uin) - uin_dn
    return uscat_dn + 1j * eta * uscat
