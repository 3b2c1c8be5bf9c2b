"""NumPy FP64 restatement of the reference's HPS hot path (CPU oracle).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import it; the ``jaxhps_b200`` package never does (its CUDA path fails loudly when
the extension is missing — there is no CPU fallback).

What it restates (all paths relative to /root/reference/src/jaxhps):

* leaf operator assembly and the DtN / ItI local solve:
  ``local_solve/_uniform_2D_DtN.py:105-280``, ``local_solve/_uniform_3D_DtN.py:13-145``,
  ``local_solve/_uniform_2D_ItI.py:120-193``;
* the oct / quad Schur-complement merges with explicit inverses and dense B, C, D:
  ``merge/_uniform_3D_DtN.py:12-541``, ``merge/_schur_complement.py:78-237,293-774``,
  ``merge/_uniform_2D_DtN.py:13-475``, ``merge/_uniform_2D_ItI.py:19-405``;
* the downward pass: ``down_pass/_uniform_3D_DtN.py:8-251``,
  ``down_pass/_uniform_2D_DtN.py:7-194``, ``down_pass/_uniform_2D_ItI.py:8-197``.

Third-party arithmetic: the reference's ``jnp.linalg.inv`` / ``@`` live in jax/jaxlib
(unpinned ``jax>=0.4``, absent from this image); here they are ``numpy.linalg.inv`` and
``@`` (LAPACK getrf+getri / OpenBLAS).

Pinning: JAX cannot be installed here, so the reference is executed through the NumPy
``jax`` shim in ``tests/golden/jaxshim`` (the reference's own Python — index maps, block
assembly, stage loops — with NumPy arithmetic).  ``tests/golden/make_golden.py`` freezes
those outputs as fixtures and ``tests/test_oracle_golden.py`` checks this file against
them; the 2D analytic known-answer cases of the reference test-suite
(``tests/test_accuracy/cases.py``) are restated in ``tests/test_oracle_analytic.py``.

The tree topology below is derived from the children's geometric positions rather than
written as the reference's literal index lists, so that it is an independent check of the
hand-written tables the CUDA path uses.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

# =====================================================================================
# Leaf level
# =====================================================================================


def _inv(A: np.ndarray) -> np.ndarray:
    """``numpy.linalg.inv``; x87 extended-precision inputs (``numpy.longdouble`` / ``clongdouble``, which LAPACK does
    not serve) go through Gaussian elimination with partial pivoting in that precision.  Only the arbitration tests
    (tests/_longdouble.py) feed such inputs: the same restated algorithm evaluated at eps = 1.1e-19 tells which of two
    FP64 results that differ by more than the 1e-10 bar is the accurate one."""
    if A.dtype not in (np.longdouble, np.clongdouble):
        return np.linalg.inv(A)
    n = A.shape[0]
    M = np.concatenate([A, np.eye(n, dtype=A.dtype)], axis=1)
    for k in range(n):
        piv = k + int(np.argmax(np.abs(M[k:, k])))
        if piv != k:
            M[[k, piv]] = M[[piv, k]]
        M[k] = M[k] / M[k, k]
        others = np.arange(n) != k
        M[others] -= np.outer(M[others, k], M[k])
    return M[:, n:]


_COEFF_ORDER_3D = ("D_xx", "D_xy", "D_yy", "D_xz", "D_yz", "D_zz", "D_x", "D_y", "D_z", "I")
_COEFF_ORDER_2D = ("D_xx", "D_xy", "D_yy", "D_x", "D_y", "I")


def gather_coeffs(pde_problem, order: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Stack the non-None coefficient arrays in the fixed order and say which were given
    (``_gather_coeffs_3D`` `_uniform_3D_DtN.py:109-145`; ``_gather_coeffs_2D``
    `_uniform_2D_DtN.py:105-133`)."""
    arrs = [getattr(pde_problem, f"{name}_coefficients", None) for name in order]
    which = np.array([a is not None for a in arrs])
    return np.array([np.asarray(a) for a in arrs if a is not None]), which


def assemble_diff_operator(coeffs: np.ndarray, which: np.ndarray, diff_ops: Sequence[np.ndarray]) -> np.ndarray:
    """``A = sum_k diag(c_k) D_k`` over the coefficients that were given; dtype follows the
    coefficients (`_uniform_2D_DtN.py:182-218`)."""
    out = np.zeros(diff_ops[0].shape, dtype=coeffs.dtype)
    counter = 0
    for k, present in enumerate(which):
        if present:
            out = out + diff_ops[k] * coeffs[counter][:, None]
            counter += 1
    return out


def get_DtN(source: np.ndarray, A: np.ndarray, Q: np.ndarray, P: np.ndarray):
    """One leaf: explicit interior inverse, then Y, T, v, h (`_uniform_2D_DtN.py:228-273`)."""
    nb = P.shape[0]
    A_ii_inv = np.linalg.inv(A[nb:, nb:])
    A_ie = A[nb:, :nb]
    L2 = np.zeros((A.shape[0], nb), dtype=A.dtype)
    L2[:nb] = np.eye(nb)
    L2[nb:] = -1 * A_ii_inv @ A_ie
    Y = L2 @ P
    T = Q @ Y
    v = np.zeros((A.shape[0], source.shape[-1]), dtype=A.dtype)
    v[nb:] = A_ii_inv @ source[nb:]
    h = Q @ v
    return Y, T, v, h


def get_ItI(source: np.ndarray, A: np.ndarray, P: np.ndarray, QH: np.ndarray, G: np.ndarray):
    """One leaf, impedance formulation: ``B = [G; A_interior_rows]`` inverted whole
    (`_uniform_2D_ItI.py:120-186`).  Returns (R, Y, h, v) like the reference's ``get_ItI``."""
    nb = P.shape[0]
    B = np.concatenate([G, A[nb:].astype(np.complex128)], axis=0)
    B_inv = np.linalg.inv(B)
    Y = B_inv[:, :nb] @ P
    Phi = B_inv[:, nb:]
    v = Phi @ source[nb:]
    h = QH @ v
    R = QH @ Y
    return R, Y, h, v


def _leaf_operators(pde_problem, order):
    eye = np.eye(pde_problem.D_x.shape[0])
    return [eye if name == "I" else getattr(pde_problem, name) for name in order]


def local_solve_stage_uniform_3D_DtN(pde_problem):
    """(Y, T, v, h) for every leaf (`local_solve/_uniform_3D_DtN.py:13-106`)."""
    return _local_solve_DtN(pde_problem, _COEFF_ORDER_3D)


def local_solve_stage_uniform_2D_DtN(pde_problem):
    """(Y, T, v, h) for every leaf (`local_solve/_uniform_2D_DtN.py:9-102`)."""
    return _local_solve_DtN(pde_problem, _COEFF_ORDER_2D)


def _local_solve_DtN(pde_problem, order):
    coeffs, which = gather_coeffs(pde_problem, order)
    ops = _leaf_operators(pde_problem, order)
    src = np.asarray(pde_problem.source)
    multi = src.ndim == 3
    if not multi:
        src = src[..., None]
    Ys, Ts, vs, hs = [], [], [], []
    for leaf in range(src.shape[0]):
        A = assemble_diff_operator(coeffs[:, leaf], which, ops)
        Y, T, v, h = get_DtN(src[leaf], A, pde_problem.Q, pde_problem.P)
        Ys.append(Y), Ts.append(T), vs.append(v), hs.append(h)
    Y, T, v, h = (np.stack(x) for x in (Ys, Ts, vs, hs))
    if not multi:
        v, h = v[..., 0], h[..., 0]
    return Y, T, v, h


def local_solve_stage_uniform_2D_ItI(pde_problem):
    """(Y, R, v, h), complex128 (`local_solve/_uniform_2D_ItI.py:10-117`)."""
    coeffs, which = gather_coeffs(pde_problem, _COEFF_ORDER_2D)
    ops = _leaf_operators(pde_problem, _COEFF_ORDER_2D)
    src = np.asarray(pde_problem.source)
    multi = src.ndim == 3
    if not multi:
        src = src[..., None]
    Rs, Ys, hs, vs = [], [], [], []
    for leaf in range(src.shape[0]):
        A = assemble_diff_operator(coeffs[:, leaf], which, ops)
        R, Y, h, v = get_ItI(src[leaf], A, pde_problem.P, pde_problem.QH, pde_problem.G)
        Rs.append(R), Ys.append(Y), hs.append(h), vs.append(v)
    R, Y, h, v = (np.stack(x) for x in (Rs, Ys, hs, vs))
    if not multi:
        v, h = v[..., 0], h[..., 0]
    return Y, R, v, h


# =====================================================================================
# Schur complement core (`merge/_schur_complement.py:117-147, 182-237`)
# =====================================================================================


def assemble_merge_outputs(A_lst, B, C, D_inv, h_ext, h_int):
    """``S=-D^-1 C``, ``T=A (+) B S``, ``g~=-D^-1 h_int``, ``h=h_ext+B g~``."""
    S = -1 * D_inv @ C
    T = B @ S
    at = 0
    for A in A_lst:
        n = A.shape[0]
        T[at : at + n, at : at + n] += A
        at += n
    g_tilde = -1 * D_inv @ h_int
    h_out = h_ext + B @ g_tilde
    return T, S, h_out, g_tilde


# =====================================================================================
# 3D oct merge.  Children a..h sit at (ix,iy,iz); faces 0..5 = x-,x+,y-,y+,z-,z+.
# =====================================================================================

_OCT_POS = [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1), (0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]
# interface order of the interior unknowns: 9:a|b 10:b|c 11:c|d 12:d|a 13:e|f 14:f|g
# 15:g|h 16:h|e 17:a|e 18:b|f 19:c|g 20:d|h (`merge/_uniform_3D_DtN.py:238-380`)
_OCT_INTERFACES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def _oct_face_roles():
    """roles[child][face] = ("ext", None) or ("int", interface slot 0..11)."""
    roles = []
    for c, pos in enumerate(_OCT_POS):
        r = []
        for face in range(6):
            axis, side = divmod(face, 2)
            if pos[axis] == side:
                r.append(("ext", None))
            else:
                nb_pos = list(pos)
                nb_pos[axis] = side
                nb = _OCT_POS.index(tuple(nb_pos))
                slot = [i for i, pr in enumerate(_OCT_INTERFACES) if set(pr) == {c, nb}][0]
                r.append(("int", slot))
        roles.append(r)
    return roles


def _oct_parent_face_children():
    """For each parent face, the four children touching it in quad-recursion order
    SW,SE,NE,NW of the face's two free coordinates (`merge/_uniform_3D_DtN.py:507-541`)."""
    out = []
    for face in range(6):
        axis, side = divmod(face, 2)
        free = [a for a in range(3) if a != axis]
        order = []
        for u, v in ((0, 0), (1, 0), (1, 1), (0, 1)):
            pos = [0, 0, 0]
            pos[axis] = side
            pos[free[0]], pos[free[1]] = u, v
            order.append(_OCT_POS.index(tuple(pos)))
        out.append(order)
    return out


_OCT_ROLES = _oct_face_roles()
_OCT_FACE_CHILDREN = _oct_parent_face_children()


def uniform_oct_merge_DtN(T_children: np.ndarray, h_children: np.ndarray, need_T: bool = True, probe=None):
    """One oct merge as the reference performs it: dense B (24m x 12m), C, D assembled from
    the children's face blocks, explicit ``inv(D)``, then the region->face permutation
    (`merge/_uniform_3D_DtN.py:127-187`, `_schur_complement.py:293-774`).

    T_children (8, 6m, 6m); h_children (8, 6m[, n_src]).  Returns (S, T, h_out, g_tilde).

    ``need_T=False`` (fixture generation at BASELINE sizes, where the 24m x 24m ``T`` of the root
    does not fit beside its factors): ``T`` is not formed; the second return value is ``T @ probe``
    (``probe`` in the parent's face order) evaluated as ``A x (+) B (S x)``, or ``None``."""
    m = T_children.shape[-1] // 6
    tail = h_children.shape[2:]
    B = np.zeros((24 * m, 12 * m))
    C = np.zeros((12 * m, 24 * m))
    D = np.zeros((12 * m, 12 * m))
    h_int = np.zeros((12 * m,) + tail)
    h_ext = np.zeros((24 * m,) + tail)
    A_lst = []
    fs = lambda f: slice(f * m, (f + 1) * m)  # noqa: E731
    for c in range(8):
        T, h = T_children[c], h_children[c]
        ext_faces = [f for f in range(6) if _OCT_ROLES[c][f][0] == "ext"]
        int_faces = [f for f in range(6) if _OCT_ROLES[c][f][0] == "int"]
        ext_slot = {f: 3 * c + i for i, f in enumerate(ext_faces)}
        int_slot = {f: _OCT_ROLES[c][f][1] for f in int_faces}
        ext_idx = np.concatenate([np.arange(f * m, (f + 1) * m) for f in ext_faces])
        A_lst.append(T[np.ix_(ext_idx, ext_idx)])
        for f in ext_faces:
            h_ext[fs(ext_slot[f])] = h[fs(f)]
            for g in int_faces:
                B[fs(ext_slot[f]), fs(int_slot[g])] = T[fs(f), fs(g)]
                C[fs(int_slot[g]), fs(ext_slot[f])] = T[fs(g), fs(f)]
        for f in int_faces:
            h_int[fs(int_slot[f])] += h[fs(f)]
            for g in int_faces:
                D[fs(int_slot[f]), fs(int_slot[g])] += T[fs(f), fs(g)]
    D_inv = np.linalg.inv(D)
    r = oct_region_to_face_permutation(m)
    if not need_T:
        S = -1 * D_inv @ C
        g_tilde = -1 * D_inv @ h_int
        h_out = h_ext + B @ g_tilde
        Tx = None
        if probe is not None:
            x = np.zeros_like(np.asarray(probe, dtype=float))
            x[r] = probe  # face order -> region order
            Tx = B @ (S @ x)
            at = 0
            for A in A_lst:
                n = A.shape[0]
                Tx[at : at + n] += A @ x[at : at + n]
                at += n
            Tx = Tx[r]
        return S[:, r], Tx, h_out[r], g_tilde
    T, S, h_out, g_tilde = assemble_merge_outputs(A_lst, B, C, D_inv, h_ext, h_int)
    return S[:, r], T[np.ix_(r, r)], h_out[r], g_tilde


def oct_region_to_face_permutation(m: int) -> np.ndarray:
    """Index vector turning [child a ext faces, ..., child h ext faces] into
    [face 0 panels, ..., face 5 panels] (`merge/_uniform_3D_DtN.py:443-541`)."""
    out = []
    for face in range(6):
        for c in _OCT_FACE_CHILDREN[face]:
            ext_faces = [f for f in range(6) if _OCT_ROLES[c][f][0] == "ext"]
            slot = 3 * c + ext_faces.index(face)
            out.append(np.arange(slot * m, (slot + 1) * m))
    return np.concatenate(out)


def merge_stage_uniform_3D_DtN(T_arr: np.ndarray, h_arr: np.ndarray, l: int, return_T: bool = False):
    """Level loop (`merge/_uniform_3D_DtN.py:12-124`).  ``S_lst``/``g_tilde_lst`` run from
    the level above the leaves to the root; the root entries carry no batch axis."""
    S_lst: List[np.ndarray] = []
    g_lst: List[np.ndarray] = []
    for _ in range(l - 1, 0, -1):
        n = T_arr.shape[0] // 8
        outs = [uniform_oct_merge_DtN(T_arr[8 * i : 8 * i + 8], h_arr[8 * i : 8 * i + 8]) for i in range(n)]
        S_lst.append(np.stack([o[0] for o in outs]))
        T_arr = np.stack([o[1] for o in outs])
        h_arr = np.stack([o[2] for o in outs])
        g_lst.append(np.stack([o[3] for o in outs]))
    S, T, h, g = uniform_oct_merge_DtN(T_arr[:8], h_arr[:8])
    S_lst.append(S)
    g_lst.append(g)
    if return_T:
        return S_lst, g_lst, T
    return S_lst, g_lst


def propagate_down_oct_DtN(S: np.ndarray, g_ext: np.ndarray, g_tilde: np.ndarray) -> np.ndarray:
    """(8, 6m[, n_src]) child boundary data from the parent's (`down_pass/_uniform_3D_DtN.py:116-246`)."""
    m = g_ext.shape[0] // 24
    g_int = S @ g_ext + g_tilde
    kids = []
    for c in range(8):
        parts = []
        for face in range(6):
            kind, slot = _OCT_ROLES[c][face]
            if kind == "int":
                parts.append(g_int[slot * m : (slot + 1) * m])
            else:
                panel = _OCT_FACE_CHILDREN[face].index(c)
                at = face * 4 * m + panel * m
                parts.append(g_ext[at : at + m])
        kids.append(np.concatenate(parts))
    return np.stack(kids)


def down_pass_uniform_3D_DtN(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr):
    """Top-down sweep then ``u = Y g + v`` on every leaf (`down_pass/_uniform_3D_DtN.py:8-113`)."""
    S_lst = list(S_lst)
    g_tilde_lst = list(g_tilde_lst)
    S_lst[-1] = S_lst[-1][None]
    g_tilde_lst[-1] = g_tilde_lst[-1][None]
    bdry = np.asarray(boundary_data)[None]
    for level in range(len(S_lst) - 1, -1, -1):
        kids = [propagate_down_oct_DtN(S_lst[level][i], bdry[i], g_tilde_lst[level][i]) for i in range(bdry.shape[0])]
        bdry = np.concatenate(kids, axis=0)
    if bdry.ndim == 3:
        return np.einsum("ijk,ikl->ijl", Y_arr, bdry) + v_arr
    return np.einsum("ijk,ik->ij", Y_arr, bdry) + v_arr


# =====================================================================================
# 2D quad merge, DtN.  Children a..d = SW, SE, NE, NW; sides 0..3 = S, E, N, W, each walked
# counter-clockwise (S: W->E, E: S->N, N: E->W, W: N->S).
# =====================================================================================

_QUAD_POS = [(0, 0), (1, 0), (1, 1), (0, 1)]
# interfaces 5:a|b 6:b|c 7:c|d 8:d|a; the first child of each pair fixes the orientation
# (`merge/_uniform_2D_DtN.py:385-437`)
_QUAD_INTERFACES = [(0, 1), (1, 2), (2, 3), (3, 0)]
# side -> (axis normal to it, which end): S is y-, E is x+, N is y+, W is x-
_QUAD_SIDE_AXIS = [(1, 0), (0, 1), (1, 1), (0, 0)]


def _quad_roles():
    """roles[child][side] = ("ext", panel) with the parent's boundary panels numbered
    counter-clockwise from the SW corner (a.S b.S b.E c.E c.N d.N d.W a.W), or
    ("int", slot, flipped)."""
    panel_order = [(0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2), (3, 3), (0, 3)]  # (child, side)
    roles = []
    for c, pos in enumerate(_QUAD_POS):
        r = []
        for side in range(4):
            axis, end = _QUAD_SIDE_AXIS[side]
            if pos[axis] == end:
                r.append(("ext", panel_order.index((c, side)), False))
            else:
                nb_pos = list(pos)
                nb_pos[axis] = end
                nb = _QUAD_POS.index(tuple(nb_pos))
                slot = [i for i, pr in enumerate(_QUAD_INTERFACES) if set(pr) == {c, nb}][0]
                r.append(("int", slot, _QUAD_INTERFACES[slot][0] != c))
        roles.append(r)
    return roles


_QUAD_ROLES = _quad_roles()


def _side_idx(side: int, m: int, flipped: bool) -> np.ndarray:
    idx = np.arange(side * m, (side + 1) * m)
    return idx[::-1] if flipped else idx


def uniform_quad_merge_DtN(T_children: np.ndarray, h_children: np.ndarray, return_ops: bool = False):
    """One quad merge with dense B, C, D and explicit ``inv(D)``
    (`merge/_uniform_2D_DtN.py:206-348`); exterior unknowns come out in boundary order, which
    is what the reference's ``roll(-n_int)`` achieves.  T_children (4, 4m, 4m)."""
    m = T_children.shape[-1] // 4
    tail = h_children.shape[2:]
    # assemble in the reference's pre-roll order [a:(W,S) b:(S,E) c:(E,N) d:(N,W)]
    pre = [(0, 3), (0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2), (3, 3)]  # (child, side)
    B = np.zeros((8 * m, 4 * m))
    C = np.zeros((4 * m, 8 * m))
    D = np.zeros((4 * m, 4 * m))
    h_int = np.zeros((4 * m,) + tail)
    h_ext = np.zeros((8 * m,) + tail)
    A_lst = []
    blk = lambda k: slice(k * m, (k + 1) * m)  # noqa: E731
    for c in range(4):
        T, h = T_children[c], h_children[c]
        ext = [(pre.index((c, s)), _side_idx(s, m, False)) for s in range(4) if _QUAD_ROLES[c][s][0] == "ext"]
        ext.sort(key=lambda t: t[0])
        inte = [(_QUAD_ROLES[c][s][1], _side_idx(s, m, _QUAD_ROLES[c][s][2])) for s in range(4) if _QUAD_ROLES[c][s][0] == "int"]
        ext_idx = np.concatenate([ix for _, ix in ext])
        A_lst.append(T[np.ix_(ext_idx, ext_idx)])
        for k, ix in ext:
            h_ext[blk(k)] = h[ix]
            for s, jx in inte:
                B[blk(k), blk(s)] = T[np.ix_(ix, jx)]
                C[blk(s), blk(k)] = T[np.ix_(jx, ix)]
        for s, ix in inte:
            h_int[blk(s)] += h[ix]
            for s2, jx in inte:
                D[blk(s), blk(s2)] += T[np.ix_(ix, jx)]
    D_inv = np.linalg.inv(D)
    T, S, h_out, g_tilde = assemble_merge_outputs(A_lst, B, C, D_inv, h_ext, h_int)
    T = np.roll(np.roll(T, -m, axis=0), -m, axis=1)
    if return_ops:  # the no-source build keeps D^-1 and B D^-1 (`_schur_complement.py:240-290`)
        return np.roll(S, -m, axis=1), T, D_inv, B @ D_inv
    return np.roll(S, -m, axis=1), T, np.roll(h_out, -m, axis=0), g_tilde


def merge_stage_uniform_2D_DtN(T_arr: np.ndarray, h_arr: np.ndarray, l: int, return_T: bool = False):
    """Level loop (`merge/_uniform_2D_DtN.py:13-203`); every entry of the lists keeps its batch
    axis, the root's being 1 (`:192-197`)."""
    S_lst, g_lst = [], []
    if l <= 0 and T_arr.shape[0] == 4:
        l = 1  # the reference runs l - 1 batched levels plus one final merge, also for l = 0 (`:95,141`)
    for _ in range(l):
        n = T_arr.shape[0] // 4
        outs = [uniform_quad_merge_DtN(T_arr[4 * i : 4 * i + 4], h_arr[4 * i : 4 * i + 4]) for i in range(n)]
        S_lst.append(np.stack([o[0] for o in outs]))
        T_arr = np.stack([o[1] for o in outs])
        h_arr = np.stack([o[2] for o in outs])
        g_lst.append(np.stack([o[3] for o in outs]))
    if return_T:
        return S_lst, g_lst, T_arr[0]
    return S_lst, g_lst


def propagate_down_quad_DtN(S: np.ndarray, g_ext: np.ndarray, g_tilde: np.ndarray) -> np.ndarray:
    """(4, 4m[, n_src]) child boundary data (`down_pass/_uniform_2D_DtN.py:125-189`)."""
    m = g_ext.shape[0] // 8
    g_int = S @ g_ext + g_tilde
    kids = []
    for c in range(4):
        parts = []
        for side in range(4):
            kind, k, flipped = _QUAD_ROLES[c][side]
            seg = g_ext[k * m : (k + 1) * m] if kind == "ext" else g_int[k * m : (k + 1) * m]
            parts.append(seg[::-1] if flipped else seg)
        kids.append(np.concatenate(parts))
    return np.stack(kids)


def down_pass_uniform_2D_DtN(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr):
    """(`down_pass/_uniform_2D_DtN.py:7-122`); ``Y_arr=None`` returns the leaves' boundary data."""
    bdry = np.asarray(boundary_data)[None]
    for level in range(len(S_lst) - 1, -1, -1):
        kids = [propagate_down_quad_DtN(S_lst[level][i], bdry[i], g_tilde_lst[level][i]) for i in range(bdry.shape[0])]
        bdry = np.concatenate(kids, axis=0)
    if Y_arr is None:
        return bdry
    if bdry.ndim == 3:
        return np.einsum("ijk,ikl->ijl", Y_arr, bdry) + v_arr
    return np.einsum("ijk,ik->ij", Y_arr, bdry) + v_arr


# =====================================================================================
# 2D quad merge, ItI (impedance-to-impedance, complex128).
# Every interface carries TWO unknown vectors: the incoming impedance data of each of the two
# children that share it.  Reference: `merge/_uniform_2D_ItI.py:182-375`,
# `merge/_schur_complement.py:6-41, 78-114`.
# =====================================================================================

# unknown order used while solving: [a5, a8, c6, c7 | b5, b6, d7, d8] as (child, interface slot)
_ITI_SOLVE_ORDER = [(0, 0), (0, 3), (2, 1), (2, 2), (1, 0), (1, 1), (3, 2), (3, 3)]
# order of the rows of S / g_tilde the reference returns: [a5, b5, b6, c6, c7, d7, d8, a8]
_ITI_OUT_ORDER = [(0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2), (3, 3), (0, 3)]


def invert_D_ItI(D_12: np.ndarray, D_21: np.ndarray) -> np.ndarray:
    """``(I + [[0, D12],[D21, 0]])^-1`` through the Schur complement ``W = I - D12 D21``
    (`_schur_complement.py:78-114`)."""
    n, m = D_12.shape[0], D_21.shape[0]
    W_inv = _inv(np.eye(n) - D_12 @ D_21)
    out = np.zeros((n + m, n + m), dtype=D_12.dtype)
    out[:n, :n] = W_inv
    out[:n, n:] = -1 * W_inv @ D_12
    out[n:, :n] = -1 * D_21 @ W_inv
    out[n:, n:] = np.eye(m) + D_21 @ W_inv @ D_12
    return out


def uniform_quad_merge_ItI(R_children: np.ndarray, h_children: np.ndarray, return_ops: bool = False):
    """One ItI quad merge.  R_children (4, 4m, 4m) complex; h_children (4, 4m, n_src).
    Returns (S, R, h_out, g_tilde) with S (8m, 8m), g_tilde (8m, n_src)."""
    m = R_children.shape[-1] // 4
    n_src = h_children.shape[-1]
    dt = np.result_type(R_children.dtype, np.complex128)  # complex128; clongdouble for the arbitration tests
    pre = [(0, 3), (0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2), (3, 3)]  # ext panels before the roll
    blk = lambda k: slice(k * m, (k + 1) * m)  # noqa: E731
    B = np.zeros((8 * m, 8 * m), dtype=dt)
    C = np.zeros((8 * m, 8 * m), dtype=dt)
    Dc = np.zeros((8 * m, 8 * m), dtype=dt)  # the coupling part of D (D = I + Dc)
    h_int = np.zeros((8 * m, n_src), dtype=dt)
    h_ext = np.zeros((8 * m, n_src), dtype=dt)
    A_lst = []
    pos = {key: i for i, key in enumerate(_ITI_SOLVE_ORDER)}
    for c in range(4):
        R, h = R_children[c], h_children[c]
        ext = sorted((pre.index((c, s)), _side_idx(s, m, False)) for s in range(4) if _QUAD_ROLES[c][s][0] == "ext")
        inte = [(_QUAD_ROLES[c][s][1], _side_idx(s, m, _QUAD_ROLES[c][s][2])) for s in range(4) if _QUAD_ROLES[c][s][0] == "int"]
        ext_idx = np.concatenate([ix for _, ix in ext])
        A_lst.append(R[np.ix_(ext_idx, ext_idx)])
        for k, ix in ext:
            h_ext[blk(k)] = h[ix]
            for s, jx in inte:
                B[blk(k), blk(pos[(c, s)])] = R[np.ix_(ix, jx)]  # own incoming interface data -> own exterior
        for s, ix in inte:
            # the equation for the OTHER child's unknown on interface s is driven by this child's outgoing data
            other = [y for y in _QUAD_INTERFACES[s] if y != c][0]
            row = pos[(other, s)]
            h_int[blk(row)] = h[ix]
            for k, jx in ext:
                C[blk(row), blk(k)] = R[np.ix_(ix, jx)]
            for s2, jx in inte:
                Dc[blk(row), blk(pos[(c, s2)])] = R[np.ix_(ix, jx)]
    half = 4 * m
    assert not Dc[:half, :half].any() and not Dc[half:, half:].any()
    D_inv = invert_D_ItI(Dc[:half, half:], Dc[half:, :half])
    T, S, h_out, g_tilde = assemble_merge_outputs(A_lst, B, C, D_inv, h_ext, h_int)
    T = np.roll(np.roll(T, -m, axis=0), -m, axis=1)
    S = np.roll(S, -m, axis=1)
    h_out = np.roll(h_out, -m, axis=0)
    r = np.concatenate([np.arange(pos[key] * m, (pos[key] + 1) * m) for key in _ITI_OUT_ORDER])
    if return_ops:  # D^-1 and B D^-1 stay in the solve order / pre-roll rows (`_nosource_uniform_2D_ItI.py:293-322`)
        return S[r], T, D_inv, B @ D_inv
    return S[r], T, h_out, g_tilde[r]


def merge_stage_uniform_2D_ItI(T_arr, h_arr, l: int, return_T: bool = False):
    """Level loop (`merge/_uniform_2D_ItI.py:19-179`): lists keep their batch axis (root: 1);
    single-source ``g_tilde`` entries are squeezed to (n, 8m)."""
    multi = h_arr.ndim == 3
    if not multi:
        h_arr = h_arr[..., None]
    S_lst, g_lst = [], []
    if l <= 0 and T_arr.shape[0] == 4:
        l = 1  # l - 1 batched levels plus one final merge in the reference, also for l = 0
    for _ in range(l):
        n = T_arr.shape[0] // 4
        outs = [uniform_quad_merge_ItI(T_arr[4 * i : 4 * i + 4], h_arr[4 * i : 4 * i + 4]) for i in range(n)]
        S_lst.append(np.stack([o[0] for o in outs]))
        T_arr = np.stack([o[1] for o in outs])
        h_arr = np.stack([o[2] for o in outs])
        g = np.stack([o[3] for o in outs])
        g_lst.append(g if multi else g[..., 0])
    if return_T:
        return S_lst, g_lst, T_arr[0]
    return S_lst, g_lst


def propagate_down_quad_ItI(S: np.ndarray, g_ext: np.ndarray, g_tilde: np.ndarray) -> np.ndarray:
    """(4, 4m[, n_src]) incoming impedance data of the children
    (`down_pass/_uniform_2D_ItI.py:121-192`)."""
    m = g_ext.shape[0] // 8
    t_int = S @ g_ext + g_tilde
    where = {key: i for i, key in enumerate(_ITI_OUT_ORDER)}
    kids = []
    for c in range(4):
        parts = []
        for side in range(4):
            kind, k, flipped = _QUAD_ROLES[c][side]
            if kind == "ext":
                seg = g_ext[k * m : (k + 1) * m]
            else:
                w = where[(c, k)]
                seg = t_int[w * m : (w + 1) * m]
            parts.append(seg[::-1] if flipped else seg)
        kids.append(np.concatenate(parts))
    return np.stack(kids)


def down_pass_uniform_2D_ItI(boundary_data, S_lst, g_tilde_lst, Y_arr, v_arr):
    """(`down_pass/_uniform_2D_ItI.py:8-118`)."""
    bdry = np.asarray(boundary_data)[None]
    for level in range(len(S_lst) - 1, -1, -1):
        kids = [propagate_down_quad_ItI(S_lst[level][i], bdry[i], g_tilde_lst[level][i]) for i in range(bdry.shape[0])]
        bdry = np.concatenate(kids, axis=0)
    if Y_arr is None:
        return bdry
    if bdry.ndim == 3:
        return np.einsum("ijk,ikl->ijl", Y_arr, bdry) + v_arr
    return np.einsum("ijk,ik->ij", Y_arr, bdry) + v_arr


# =====================================================================================
# Source given at solve time (2D uniform): no-source build + upward pass.
# Reference: local_solve/_nosource_uniform_2D_{DtN,ItI}.py, merge/_nosource_uniform_2D_{DtN,ItI}.py,
# up_pass/_uniform_2D_{DtN,ItI}.py, _build_solver.py:261-331, _solve.py:115-151.
# =====================================================================================


def nosource_local_solve_stage_uniform_2D_DtN(pde_problem):
    """(Y, T, Phi) with Phi = A_ii^-1, shape (n, n_i, n_i) (`_nosource_uniform_2D_DtN.py:76-123`)."""
    coeffs, which = gather_coeffs(pde_problem, _COEFF_ORDER_2D)
    ops = _leaf_operators(pde_problem, _COEFF_ORDER_2D)
    nb = pde_problem.P.shape[0]
    Ys, Ts, Phis = [], [], []
    for leaf in range(coeffs.shape[1]):
        A = assemble_diff_operator(coeffs[:, leaf], which, ops)
        Y, T, _, _ = get_DtN(np.zeros((A.shape[0], 1)), A, pde_problem.Q, pde_problem.P)
        Ys.append(Y), Ts.append(T), Phis.append(np.linalg.inv(A[nb:, nb:]))
    return np.stack(Ys), np.stack(Ts), np.stack(Phis)


def nosource_local_solve_stage_uniform_2D_ItI(pde_problem):
    """(Y, R, Phi) with Phi = B^-1[:, n_b:], shape (n, p^2, n_i) (`_nosource_uniform_2D_ItI.py:80-146`)."""
    coeffs, which = gather_coeffs(pde_problem, _COEFF_ORDER_2D)
    ops = _leaf_operators(pde_problem, _COEFF_ORDER_2D)
    nb = pde_problem.P.shape[0]
    Ys, Rs, Phis = [], [], []
    for leaf in range(coeffs.shape[1]):
        A = assemble_diff_operator(coeffs[:, leaf], which, ops)
        B_inv = np.linalg.inv(np.concatenate([pde_problem.G, A[nb:].astype(np.complex128)], axis=0))
        Y = B_inv[:, :nb] @ pde_problem.P
        Ys.append(Y), Rs.append(pde_problem.QH @ Y), Phis.append(B_inv[:, nb:])
    return np.stack(Ys), np.stack(Rs), np.stack(Phis)


def _nosource_merge_stage(T_arr, l, merge_fn, return_T):
    S_lst, D_inv_lst, BD_inv_lst = [], [], []
    h_dummy = np.zeros(T_arr.shape[:2] + (1,), dtype=T_arr.dtype)
    if l <= 0 and T_arr.shape[0] == 4:
        l = 1
    for _ in range(l):
        n = T_arr.shape[0] // 4
        outs = [merge_fn(T_arr[4 * i : 4 * i + 4], h_dummy[4 * i : 4 * i + 4], return_ops=True) for i in range(n)]
        S_lst.append(np.stack([o[0] for o in outs]))
        T_arr = np.stack([o[1] for o in outs])
        D_inv_lst.append(np.stack([o[2] for o in outs]))
        BD_inv_lst.append(np.stack([o[3] for o in outs]))
        h_dummy = np.zeros(T_arr.shape[:2] + (1,), dtype=T_arr.dtype)
    if return_T:
        return S_lst, D_inv_lst, BD_inv_lst, T_arr[0]
    return S_lst, D_inv_lst, BD_inv_lst


def nosource_merge_stage_uniform_2D_DtN(T_arr, l: int, return_T: bool = False):
    """(S_lst, D_inv_lst, BD_inv_lst[, T_last]) (`merge/_nosource_uniform_2D_DtN.py:13-130`)."""
    return _nosource_merge_stage(T_arr, l, uniform_quad_merge_DtN, return_T)


def nosource_merge_stage_uniform_2D_ItI(T_arr, l: int, return_T: bool = False):
    """(`merge/_nosource_uniform_2D_ItI.py:19-150`)."""
    return _nosource_merge_stage(T_arr, l, uniform_quad_merge_ItI, return_T)


def _children_h_parts(h4: np.ndarray, m: int):
    """Exterior (pre-roll order) and per-(child, interface) pieces of four children's h."""
    pre = [(0, 3), (0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2), (3, 3)]
    ext = np.concatenate([h4[c][_side_idx(s, m, False)] for c, s in pre])
    part = {}
    for c in range(4):
        for s in range(4):
            kind, k, flipped = _QUAD_ROLES[c][s]
            if kind == "int":
                part[(c, k)] = h4[c][_side_idx(s, m, flipped)]
    return ext, part


def up_pass_uniform_2D_DtN(source, pde_problem, return_h_last: bool = False):
    """v = [0; Phi f_i], h = Q v, then per level g~ = -D^-1 h_int, h = roll(h_ext - B D^-1 h_int)
    (`up_pass/_uniform_2D_DtN.py:8-173`).  Outputs keep the source axis (the reference does not
    squeeze in the DtN up pass)."""
    src = np.asarray(source)
    if src.ndim == 2:
        src = src[..., None]
    nb = pde_problem.P.shape[0]
    v = np.zeros_like(src)
    v[:, nb:] = np.einsum("ijk,ikl->ijl", pde_problem.Phi, src[:, nb:])
    h = np.einsum("ij,kjl->kil", pde_problem.Q, v)
    g_lst = []
    for D_inv, BD_inv in zip(pde_problem.D_inv_lst, pde_problem.BD_inv_lst):
        m = h.shape[1] // 4
        new_h, g = [], []
        for i in range(h.shape[0] // 4):
            ext, part = _children_h_parts(h[4 * i : 4 * i + 4], m)
            h_int = np.concatenate([part[(a, s)] + part[(b, s)] for s, (a, b) in enumerate(_QUAD_INTERFACES)])
            g.append(-1 * D_inv[i] @ h_int)
            new_h.append(np.roll(ext - BD_inv[i] @ h_int, -m, axis=0))
        h = np.stack(new_h)
        g_lst.append(np.stack(g))
    if return_h_last:
        return v, g_lst, h[0]
    return v, g_lst


def up_pass_uniform_2D_ItI(source, pde_problem, return_h_last: bool = False):
    """(`up_pass/_uniform_2D_ItI.py:8-219`): h_int lists the OTHER child's outgoing data in the solve
    order, g~ is returned in the out order; single-source outputs are squeezed."""
    src = np.asarray(source)
    multi = src.ndim == 3
    if not multi:
        src = src[..., None]
    nb = pde_problem.P.shape[0]
    v = np.einsum("ijk,ikl->ijl", pde_problem.Phi, src[:, nb:])
    h = np.einsum("ij,kjl->kil", pde_problem.QH, v)
    pos = {key: i for i, key in enumerate(_ITI_SOLVE_ORDER)}
    g_lst = []
    for D_inv, BD_inv in zip(pde_problem.D_inv_lst, pde_problem.BD_inv_lst):
        m = h.shape[1] // 4
        r = np.concatenate([np.arange(pos[key] * m, (pos[key] + 1) * m) for key in _ITI_OUT_ORDER])
        new_h, g = [], []
        for i in range(h.shape[0] // 4):
            ext, part = _children_h_parts(h[4 * i : 4 * i + 4], m)
            rows = []
            for (c, s) in _ITI_SOLVE_ORDER:
                other = [y for y in _QUAD_INTERFACES[s] if y != c][0]
                rows.append(part[(other, s)])
            h_int = np.concatenate(rows)
            g.append((-1 * D_inv[i] @ h_int)[r])
            new_h.append(np.roll(ext - BD_inv[i] @ h_int, -m, axis=0))
        h = np.stack(new_h)
        g_lst.append(np.stack(g))
    if not multi:
        v, g_lst, h = v[..., 0], [g[..., 0] for g in g_lst], h[..., 0]
    if return_h_last:
        return v, g_lst, h[0]
    return v, g_lst
