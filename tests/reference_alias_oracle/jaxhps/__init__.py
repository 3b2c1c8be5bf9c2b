"""TEST INFRASTRUCTURE: `jaxhps` alias whose STAGE functions are the CPU oracle (oracle/hps_oracle.py) and whose
host layer is jaxhps_b200's — used to run the reference's own known-answer accuracy suite
(/root/reference/tests/test_accuracy) directly against the oracle, on the CPU."""
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from jaxhps_b200 import Domain, DiscretizationNode2D, DiscretizationNode3D, PDEProblem  # noqa: E402,F401
from jaxhps_b200 import _domain, _pdeproblem, _tree  # noqa: E402
from oracle import hps_oracle as _o  # noqa: E402


def _mod(name, **attrs):
    m = types.ModuleType("jaxhps." + name)
    m.__dict__.update(attrs)
    sys.modules["jaxhps." + name] = m
    return m


def _kw(fn, *names):
    """Accept the reference's keyword names / device arguments and forward positionally."""
    def wrapped(*args, device=None, host_device=None, **kw):
        args = list(args) + [kw.pop(n) for n in names[len(args):] if n in kw]
        return fn(*args, **kw)

    return wrapped


sys.modules["jaxhps._domain"] = _domain
sys.modules["jaxhps._pdeproblem"] = _pdeproblem
sys.modules["jaxhps._discretization_tree"] = _tree

_ls = dict(
    local_solve_stage_uniform_2D_DtN=_kw(_o.local_solve_stage_uniform_2D_DtN, "pde_problem"),
    local_solve_stage_uniform_2D_ItI=_kw(_o.local_solve_stage_uniform_2D_ItI, "pde_problem"),
    local_solve_stage_uniform_3D_DtN=_kw(_o.local_solve_stage_uniform_3D_DtN, "pde_problem"),
    nosource_local_solve_stage_uniform_2D_DtN=_kw(_o.nosource_local_solve_stage_uniform_2D_DtN, "pde_problem"),
    nosource_local_solve_stage_uniform_2D_ItI=_kw(_o.nosource_local_solve_stage_uniform_2D_ItI, "pde_problem"),
)
_mg = dict(
    merge_stage_uniform_2D_DtN=_kw(_o.merge_stage_uniform_2D_DtN, "T_arr", "h_arr", "l"),
    merge_stage_uniform_2D_ItI=_kw(_o.merge_stage_uniform_2D_ItI, "T_arr", "h_arr", "l"),
    merge_stage_uniform_3D_DtN=_kw(_o.merge_stage_uniform_3D_DtN, "T_arr", "h_arr", "l"),
    nosource_merge_stage_uniform_2D_DtN=_kw(_o.nosource_merge_stage_uniform_2D_DtN, "T_arr", "l"),
    nosource_merge_stage_uniform_2D_ItI=_kw(_o.nosource_merge_stage_uniform_2D_ItI, "T_arr", "l"),
)
_dp = dict(
    down_pass_uniform_2D_DtN=_kw(_o.down_pass_uniform_2D_DtN, "boundary_data", "S_lst", "g_tilde_lst", "Y_arr", "v_arr"),
    down_pass_uniform_2D_ItI=_kw(_o.down_pass_uniform_2D_ItI, "boundary_data", "S_lst", "g_tilde_lst", "Y_arr", "v_arr"),
    down_pass_uniform_3D_DtN=_kw(_o.down_pass_uniform_3D_DtN, "boundary_data", "S_lst", "g_tilde_lst", "Y_arr", "v_arr"),
)
_up = dict(
    up_pass_uniform_2D_DtN=_kw(_o.up_pass_uniform_2D_DtN, "source", "pde_problem"),
    up_pass_uniform_2D_ItI=_kw(_o.up_pass_uniform_2D_ItI, "source", "pde_problem"),
)
local_solve = _mod("local_solve", **_ls)
merge = _mod("merge", **_mg)
down_pass = _mod("down_pass", **_dp)
up_pass = _mod("up_pass", **_up)
for _n in ("_uniform_2D_DtN", "_uniform_2D_ItI", "_uniform_3D_DtN", "_nosource_uniform_2D_DtN", "_nosource_uniform_2D_ItI"):
    _mod("local_solve." + _n, **_ls)
    _mod("merge." + _n, **_mg)
    _mod("down_pass." + _n, **_dp)
    _mod("up_pass." + _n, **_up)
