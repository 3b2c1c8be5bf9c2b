"""NumPy FP64 restatement of the reference's ADAPTIVE (non-uniform tree) DtN path — CPU oracle.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as ``oracle/hps_oracle.py``: only
``tests/``, ``smoke()`` and the CPU-baseline legs of ``bench.py`` may import it).

What it restates (paths relative to /root/reference/src/jaxhps):

* per-leaf rescaling of the unit-box operators by the leaf's side length and the DtN local solve:
  ``local_solve/_adaptive_2D_DtN.py:12-172``, ``local_solve/_adaptive_3D_DtN.py:13-175``;
* which interface panels must be coarsened (a leaf facing four / two half-size leaves):
  ``merge/_utils_adaptive_3D_DtN.py:11-176`` (``_projection_lst``), ``merge/_utils_adaptive_2D_DtN.py:8-165``;
* the merge of a node's children with projected interface blocks — gather per-child blocks,
  coarsen rows with ``L_1f4``/``L_1f2`` and columns with ``L_4f1``/``L_2f1``, dense B, C, D, explicit
  ``inv(D)``, exterior unknowns permuted to boundary order: ``merge/_adaptive_3D_DtN.py:150-347``,
  ``merge/_utils_adaptive_3D_DtN.py:179-881``, ``merge/_adaptive_2D_DtN.py:160-433``,
  ``merge/_utils_adaptive_2D_DtN.py:168-584``, ``merge/_schur_complement.py:117-237``;
* the level loops ``merge/_adaptive_3D_DtN.py:30-147`` / ``merge/_adaptive_2D_DtN.py:24-93``;
* the downward pass with re-refinement of coarsened interface data:
  ``down_pass/_adaptive_3D_DtN.py:15-399``, ``down_pass/_adaptive_2D_DtN.py:13-263``.

Pinning: ``tests/golden/make_golden_adaptive.py`` executes the unmodified reference on the NumPy
``jax`` shim for six adaptive trees (generated and hand-made, 2D and 3D) and
``tests/test_oracle_adaptive.py`` checks this file against those fixtures.

Per-node results live in a dict keyed by ``id(node)`` instead of ``node.data`` so that the oracle and
the product can work on the same tree without seeing each other's arrays.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .hps_oracle import (
    _COEFF_ORDER_2D,
    _COEFF_ORDER_3D,
    _OCT_FACE_CHILDREN,
    _OCT_ROLES,
    _QUAD_ROLES,
    assemble_diff_operator,
    assemble_merge_outputs,
    gather_coeffs,
    get_DtN,
)

# parent side -> the two children touching it, walking counter-clockwise (2D)
_QUAD_SIDE_CHILDREN = [[0, 1], [1, 2], [2, 3], [3, 0]]


def _dim(node) -> int:
    return 3 if hasattr(node, "zmin") else 2


def _leaves(node) -> List:
    if not len(node.children):
        return [node]
    return [leaf for c in node.children for leaf in _leaves(c)]


def _depth(node) -> int:
    return max(leaf.depth for leaf in _leaves(node))


def _nodes_at_level(node, level: int) -> List:
    if node.depth == level:
        return [node]
    return [n for c in node.children for n in _nodes_at_level(c, level)]


# ------------------------------------------------------------------------------- leaf stage


def local_solve_stage_adaptive_DtN(pde_problem):
    """(Y, T, v, h) per leaf.  The unit-box differentiation operators are divided by the leaf's half
    side length (squared for second derivatives) and Q is rebuilt from the scaled first derivatives
    (`local_solve/_adaptive_3D_DtN.py:128-167`, `local_solve/_adaptive_2D_DtN.py:127-164`)."""
    from jaxhps_b200._operators import precompute_Q_2D_DtN, precompute_Q_3D_DtN  # host-side constants only

    dom = pde_problem.domain
    two_d = dom.bool_2D
    order = _COEFF_ORDER_2D if two_d else _COEFF_ORDER_3D
    n_second = 3 if two_d else 6
    coeffs, which = gather_coeffs(pde_problem, order)
    eye = np.eye(pde_problem.D_x.shape[0])
    unit_ops = [eye if name == "I" else getattr(pde_problem, name) for name in order]
    src = np.asarray(pde_problem.source)[..., None]
    out = [[], [], [], []]
    for leaf in range(src.shape[0]):
        hs = pde_problem.sidelens[leaf] / 2
        ops = [op / hs**2 if k < n_second else (op / hs if name != "I" else op)
               for k, (name, op) in enumerate(zip(order, unit_ops))]
        A = assemble_diff_operator(coeffs[:, leaf], which, ops)
        first = ops[n_second : n_second + (2 if two_d else 3)]
        Q = precompute_Q_2D_DtN(dom.p, dom.q, *first) if two_d else precompute_Q_3D_DtN(dom.p, dom.q, *first)
        for lst, arr in zip(out, get_DtN(src[leaf], A, Q, pde_problem.P)):
            lst.append(arr)
    Y, T, v, h = (np.stack(x) for x in out)
    return Y, T, v[..., 0], h[..., 0]


# ------------------------------------------------------------------------------- geometry


def _face_leaves(node, face: int) -> List:
    """Leaves of ``node``'s subtree touching its face/side ``face`` in boundary-vector order.
    3D: depth-first with the children visited in the quad order of the face's plane, filtered by
    the face coordinate (`_grid_creation_3D.py:376-417`).  2D: depth-first SW,SE,NE,NW filtered by
    the side coordinate gives S and E in walking order and N, W reversed
    (`merge/_utils_adaptive_2D_DtN.py:8-143`); the counter-clockwise walk reverses those two."""
    if _dim(node) == 3:
        order = {0: [4, 7, 3, 0, 5, 6, 2, 1], 1: [4, 5, 1, 0, 7, 6, 2, 3], 2: list(range(8))}[face // 2]
        key = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[face]

        def walk(n):
            if not len(n.children):
                return [n]
            return [leaf for c in order for leaf in walk(n.children[c])]

        return [leaf for leaf in walk(node) if getattr(leaf, key) == getattr(node, key)]
    key = ("ymin", "xmax", "ymax", "xmin")[face]
    lst = [leaf for leaf in _leaves(node) if getattr(leaf, key) == getattr(node, key)]
    if face == 2:  # depth-first order meets the N side NE before NW already (children c, d)
        return lst
    if face == 3:  # W side: depth-first gives SW before NW; the walk goes north -> south
        return lst[::-1]
    return lst


def _projection_lst(lst_0: List, lst_1: List, group: int):
    """Walk two faces of an interface panel by panel; where one side has ``group`` small leaves
    against one big leaf its panels are flagged for coarsening
    (`merge/_utils_adaptive_3D_DtN.py:11-58`, group 4; `_utils_adaptive_2D_DtN.py:8-38`, group 2)."""
    out_0, out_1 = [], []
    i0 = i1 = 0
    while i0 < len(lst_0) and i1 < len(lst_1):
        len_0 = lst_0[i0].xmax - lst_0[i0].xmin
        len_1 = lst_1[i1].xmax - lst_1[i1].xmin
        if len_0 == len_1:
            out_0.append(False), out_1.append(False)
            i0 += 1
            i1 += 1
        elif len_0 < len_1:
            out_0.extend([True] * group), out_1.append(False)
            i0 += group
            i1 += 1
        else:
            out_1.extend([True] * group), out_0.append(False)
            i1 += group
            i0 += 1
    return np.array(out_0, dtype=bool), np.array(out_1, dtype=bool)


def _coarsen(M: np.ndarray, L: np.ndarray, bools, n_pp: int, group: int, axis: int) -> np.ndarray:
    """Replace every flagged run of ``group`` panels along ``axis`` by one panel: ``L @ rows`` /
    ``cols @ L`` (`merge/_utils_adaptive_3D_DtN.py:640-716`, `_utils_adaptive_2D_DtN.py:529-584`)."""
    if bools is None or not np.any(bools):
        return M
    parts, i = [], 0
    while i < len(bools):
        width = group if bools[i] else 1
        sl = [slice(None)] * M.ndim
        sl[axis] = slice(i * n_pp, (i + width) * n_pp)
        blk = M[tuple(sl)]
        if bools[i]:
            blk = L @ blk if axis == 0 else blk @ L
        parts.append(blk)
        i += width
    return np.concatenate(parts, axis=axis)


def _child_faces(dim: int, c: int):
    """[(face, "ext"|"int", slot, flipped)] of child ``c`` inside its parent."""
    if dim == 3:
        return [(f, kind, slot, False) for f, (kind, slot) in enumerate(_OCT_ROLES[c])]
    return [(f, kind, slot, flipped) for f, (kind, slot, flipped) in enumerate(_QUAD_ROLES[c])]


def _interface_bools(children, dim: int) -> Dict:
    """{(child, face): bools} for every interface face."""
    group = 4 if dim == 3 else 2
    by_slot: Dict[int, List] = {}
    for c, child in enumerate(children):
        for f, kind, slot, flipped in _child_faces(dim, c):
            if kind == "int":
                lst = _face_leaves(child, f)
                by_slot.setdefault(slot, []).append((c, f, lst[::-1] if flipped else lst))
    out = {}
    for slot, ((c0, f0, l0), (c1, f1, l1)) in by_slot.items():
        out[(c0, f0)], out[(c1, f1)] = _projection_lst(l0, l1, group)
    return out


def _face_index_ranges(child, dim: int):
    sizes = [getattr(child, f"n_{f}") for f in range(2 * dim)]
    off = np.concatenate([[0], np.cumsum(sizes)])
    return [np.arange(off[f], off[f + 1]) for f in range(2 * dim)]


# ------------------------------------------------------------------------------- merge


def adaptive_merge_node(children, T_lst, h_lst, L_refine: np.ndarray, L_coarsen: np.ndarray, q: int):
    """Merge the 4 / 8 children of one node (`merge/_adaptive_3D_DtN.py:150-347`,
    `merge/_adaptive_2D_DtN.py:160-433`).  Returns (S, T, h, g_tilde), exterior unknowns in the
    parent's boundary order."""
    dim = _dim(children[0])
    n_pp, group = (q * q, 4) if dim == 3 else (q, 2)
    bools = _interface_bools(children, dim)
    n_slots = 12 if dim == 3 else 4
    # 2D: the reference lists child a's exterior as (W, S) and rolls afterwards; we record segments
    # per (child, face) and place them directly
    blocks = []  # per child: list of (kind, key, face idx, bools)
    slot_size = {}
    for c, child in enumerate(children):
        rng = _face_index_ranges(child, dim)
        parts = []
        for f, kind, slot, flipped in _child_faces(dim, c):
            idx = rng[f][::-1] if flipped else rng[f]
            b = bools.get((c, f)) if kind == "int" else None
            n_after = len(idx) if b is None else n_pp * (len(b) - (group - 1) * int(np.sum(b)) // group)
            if kind == "int":
                assert slot_size.setdefault(slot, n_after) == n_after, "interface sizes disagree after coarsening"
            parts.append((kind, slot if kind == "int" else f, idx, b, n_after))
        blocks.append(parts)
    int_off = np.concatenate([[0], np.cumsum([slot_size[s] for s in range(n_slots)])])
    # exterior segments in final boundary order: face by face, children in the face's panel order
    face_children = _OCT_FACE_CHILDREN if dim == 3 else _QUAD_SIDE_CHILDREN
    ext_off, at = {}, 0
    for f in range(2 * dim):
        for c in face_children[f]:
            ext_off[(c, f)] = at
            at += getattr(children[c], f"n_{f}")
    n_ext, n_int = at, int(int_off[-1])
    tail = np.asarray(h_lst[0]).shape[1:]
    A = np.zeros((n_ext, n_ext))
    B = np.zeros((n_ext, n_int))
    C = np.zeros((n_int, n_ext))
    D = np.zeros((n_int, n_int))
    h_ext = np.zeros((n_ext,) + tail)
    h_int = np.zeros((n_int,) + tail)
    for c, parts in enumerate(blocks):
        T, h = np.asarray(T_lst[c]), np.asarray(h_lst[c])
        for kind_r, key_r, idx_r, b_r, n_r in parts:
            rows = T[idx_r]
            rows = _coarsen(rows, L_coarsen, b_r, n_pp, group, axis=0)
            hv = _coarsen(h[idx_r], L_coarsen, b_r, n_pp, group, axis=0)
            r0 = ext_off[(c, key_r)] if kind_r == "ext" else int_off[key_r]
            if kind_r == "ext":
                h_ext[r0 : r0 + n_r] = hv
            else:
                h_int[r0 : r0 + n_r] += hv
            for kind_c, key_c, idx_c, b_c, n_c in parts:
                blk = _coarsen(rows[:, idx_c], L_refine, b_c, n_pp, group, axis=1)
                c0 = ext_off[(c, key_c)] if kind_c == "ext" else int_off[key_c]
                target = {("ext", "ext"): A, ("ext", "int"): B, ("int", "ext"): C, ("int", "int"): D}[(kind_r, kind_c)]
                target[r0 : r0 + n_r, c0 : c0 + n_c] += blk
    D_inv = np.linalg.inv(D)
    T_out, S, h_out, g_tilde = assemble_merge_outputs([A], B, C, D_inv, h_ext, h_int)
    return S, T_out, h_out, g_tilde


def merge_stage_adaptive_DtN(pde_problem, T_arr, h_arr) -> Dict:
    """Bottom-up merge of the whole tree; returns {id(node): dict(T, h, S, g_tilde)} for every node
    (leaves carry T, h only) (`merge/_adaptive_3D_DtN.py:30-147`, `merge/_adaptive_2D_DtN.py:24-93`)."""
    dom = pde_problem.domain
    root = dom.root
    L_refine, L_coarsen = (pde_problem.L_2f1, pde_problem.L_1f2) if dom.bool_2D else (pde_problem.L_4f1, pde_problem.L_1f4)
    store: Dict[int, dict] = {}
    for i, leaf in enumerate(_leaves(root)):
        store[id(leaf)] = dict(T=np.asarray(T_arr[i]), h=np.asarray(h_arr[i]))
    for level in range(_depth(root) - 1, -1, -1):
        for node in _nodes_at_level(root, level):
            if not len(node.children):
                continue
            S, T, h, g = adaptive_merge_node(
                node.children, [store[id(c)]["T"] for c in node.children], [store[id(c)]["h"] for c in node.children],
                L_refine, L_coarsen, dom.q)
            store[id(node)] = dict(T=T, h=h, S=S, g_tilde=g)
    return store


# ------------------------------------------------------------------------------- down pass


def _refine_interface(g_seg: np.ndarray, bools, L_refine: np.ndarray, n_pp: int, group: int) -> np.ndarray:
    """Inverse walk of ``_coarsen`` on interface data: a panel that stood for ``group`` fine panels is
    interpolated back with ``L_refine`` (`down_pass/_adaptive_3D_DtN.py:132-172`)."""
    if bools is None or not np.any(bools):
        return g_seg
    parts, i, at = [], 0, 0
    while i < len(bools):
        panel = g_seg[at : at + n_pp]
        parts.append(L_refine @ panel if bools[i] else panel)
        i += group if bools[i] else 1
        at += n_pp
    return np.concatenate(parts)


def propagate_down_adaptive(node, S, g_tilde, g_ext: np.ndarray, L_refine: np.ndarray, q: int) -> List[np.ndarray]:
    """Boundary data of ``node`` (one vector in boundary order) -> boundary data of each child
    (`down_pass/_adaptive_3D_DtN.py:175-394`, `down_pass/_adaptive_2D_DtN.py:138-259`)."""
    children = node.children
    dim = _dim(node)
    n_pp, group = (q * q, 4) if dim == 3 else (q, 2)
    g_int = S @ g_ext + g_tilde
    bools = _interface_bools(children, dim)
    face_children = _OCT_FACE_CHILDREN if dim == 3 else _QUAD_SIDE_CHILDREN
    ext_off, at = {}, 0
    for f in range(2 * dim):
        for c in face_children[f]:
            ext_off[(c, f)] = at
            at += getattr(children[c], f"n_{f}")
    # interface offsets: size of each slot after coarsening
    n_slots = 12 if dim == 3 else 4
    size = {}
    for c, child in enumerate(children):
        for f, kind, slot, _ in _child_faces(dim, c):
            if kind == "int":
                b = bools[(c, f)]
                size[slot] = n_pp * (len(b) - (group - 1) * int(np.sum(b)) // group)
    int_off = np.concatenate([[0], np.cumsum([size[s] for s in range(n_slots)])])
    out = []
    for c, child in enumerate(children):
        parts = []
        for f, kind, slot, flipped in _child_faces(dim, c):
            if kind == "ext":
                parts.append(g_ext[ext_off[(c, f)] : ext_off[(c, f)] + getattr(child, f"n_{f}")])
            else:
                seg = _refine_interface(g_int[int_off[slot] : int_off[slot + 1]], bools[(c, f)], L_refine, n_pp, group)
                parts.append(seg[::-1] if flipped else seg)
        out.append(np.concatenate(parts))
    return out


def down_pass_adaptive_DtN(pde_problem, store: Dict, boundary_data, Y_arr, v_arr) -> np.ndarray:
    """``boundary_data``: list with one array per side / face of the root.  Returns (n_leaves, p^d)
    (`down_pass/_adaptive_3D_DtN.py:15-129`, `down_pass/_adaptive_2D_DtN.py:13-84`)."""
    dom = pde_problem.domain
    L_refine = pde_problem.L_2f1 if dom.bool_2D else pde_problem.L_4f1
    g_of = {id(dom.root): np.concatenate([np.asarray(b) for b in boundary_data])}
    stack = [dom.root]
    while stack:
        node = stack.pop()
        if not len(node.children):
            continue
        rec = store[id(node)]
        for child, g in zip(node.children, propagate_down_adaptive(node, rec["S"], rec["g_tilde"], g_of[id(node)], L_refine, dom.q)):
            g_of[id(child)] = g
            stack.append(child)
    return np.stack([Y_arr[i] @ g_of[id(leaf)] + v_arr[i] for i, leaf in enumerate(_leaves(dom.root))])
