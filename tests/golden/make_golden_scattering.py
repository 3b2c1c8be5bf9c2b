"""Golden vectors for the scattering coupling: executes the reference's OWN functions from
/root/reference/examples/wave_scattering_utils.py on the NumPy `jax` shim (h5py, which that module imports for its
MATLAB loader, is absent here and replaced by an empty stand-in; the loader is not used).

    python tests/golden/make_golden_scattering.py     # writes tests/golden/scattering_reference.npz"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, "/root/reference/examples")
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.modules.setdefault("h5py", types.ModuleType("h5py"))

import wave_scattering_utils as ref  # noqa: E402
from _cases import scattering_inputs  # noqa: E402

R, S, D, pts, dirs, k, eta = scattering_inputs()
T = np.asarray(ref.get_DtN_from_ItI(R, eta))
uin, normals = (np.asarray(x) for x in ref.get_uin_and_normals(k, pts, dirs))
A, b = (np.asarray(x) for x in ref.setup_scattering_lin_system(S, D, T, pts, k, dirs))
imp = np.asarray(ref.get_scattering_uscat_impedance(S, D, T, dirs, pts, k, eta))
np.savez_compressed(os.path.join(HERE, "scattering_reference.npz"), T=T, uin=uin, normals=normals, A=A, b=b, imp=imp)
print("wrote scattering_reference.npz", T.shape, imp.shape)
