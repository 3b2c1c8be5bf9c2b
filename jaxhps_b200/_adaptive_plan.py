"""Host-side "merge plan compiler" for non-uniform trees.

The reference re-derives, for every node and at every merge / down-pass call, Python lists of index
arrays and projection flags (`merge/_utils_adaptive_3D_DtN.py:61-176, 179-570, 719-881`,
`merge/_utils_adaptive_2D_DtN.py:8-165, 168-373`, `down_pass/_adaptive_3D_DtN.py:132-394`).  Here the
tree is compiled ONCE into small int32 tables at *panel* granularity (a panel = the q^(d-1) Gauss
points of one leaf face) which the CUDA kernels in ``csrc/adaptive.cu`` expand on the fly:

``seg``       per child, (n'_panels, 3) ``[start, width, rev]``: panel P of the child's interface-ready
              operator T' is the run of ``width`` panels of T starting at element ``start`` (width 1:
              copied; width ``group`` = 2^(d-1): coarsened), walked backwards if ``rev``;
``int_tbl``   (NI, 4) ``[child A, panel in A', child B, panel in B']`` per interface panel;
``ext_tbl``   (NE, 2) ``[child, panel in T']`` per exterior panel, already in the parent's boundary order;
``bs_tbl``    (host only) the non-zero blocks of B = T'[exterior face, interface face] per child, as
              [child, row0, col0, M, K, first row of S, first row of T_out] in points;
``down_tbl``  (sum n'_panels, 5) ``[child, source panel, start, width, rev]`` for the down pass, source
              panel < NE: exterior panel of the parent's data, otherwise interface panel - NE.

Nodes all of whose children are leaves are *not* planned here: they go through the batched uniform
merge kernels (the reference does the same for the deepest level only, `merge/_adaptive_3D_DtN.py:85-121`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from ._tree import FACE_CHILDREN_2D, FACE_CHILDREN_3D, _FACE_AXIS_2D, _FACE_AXIS_3D, _OFFSETS_2D, _OFFSETS_3D, _is_2D

# interface order of the interior unknowns and which child fixes the orientation (listed first):
# 3D 9:a|b 10:b|c 11:c|d 12:d|a 13:e|f 14:f|g 15:g|h 16:h|e 17:a|e 18:b|f 19:c|g 20:d|h
# (`merge/_utils_adaptive_3D_DtN.py:61-176`); 2D 5:a|b 6:b|c 7:c|d 8:d|a, the second child walks the
# interface backwards (`merge/_utils_adaptive_2D_DtN.py:168-290`).
_SLOTS_3D = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))
_SLOTS_2D = ((0, 1), (1, 2), (2, 3), (3, 0))


def _shared_faces(offsets, face_axis, a: int, b: int):
    """Faces of children ``a`` and ``b`` lying on their common interface."""
    (ax,) = [k for k in range(len(offsets[a])) if offsets[a][k] != offsets[b][k]]
    fa = face_axis.index((ax, offsets[b][ax]))  # a's face on b's side
    fb = face_axis.index((ax, offsets[a][ax]))
    return fa, fb


@dataclass
class ChildPlan:
    n: int  #: points on the child's boundary (order of T)
    n_out: int  #: points after coarsening (order of T')
    seg: np.ndarray  #: (n_out / npp, 3) int32
    identity: bool  #: T' == T (no coarsening, no reversed interface)
    face_off_out: List[int] = field(default_factory=list)  #: first T' panel of each face


@dataclass
class NodePlan:
    node: object
    npp: int
    group: int
    children: List[ChildPlan]
    int_tbl: np.ndarray
    ext_tbl: np.ndarray
    down_tbl: np.ndarray
    bs_tbl: np.ndarray  #: (n_blocks, 7) non-zero blocks of B: [child, row0, col0 in T', M, K, row0 in S, row0 in T_out]
    n_int: int
    n_ext: int
    all_leaf_children: bool
    # filled by upload(): offsets of the tables inside the tree-wide device buffer
    off: Dict[str, int] = field(default_factory=dict)


class TreePlan:
    """All per-node tables of one tree + the face panel lists they were derived from."""

    def __init__(self, root, q: int):
        self.root = root
        self.two_d = _is_2D(root)
        self.dim = 2 if self.two_d else 3
        self.npp = q if self.two_d else q * q
        self.group = 2 if self.two_d else 4
        self.n_faces = 2 * self.dim
        self._face_children = FACE_CHILDREN_2D if self.two_d else FACE_CHILDREN_3D
        self._offsets = _OFFSETS_2D if self.two_d else _OFFSETS_3D
        self._face_axis = _FACE_AXIS_2D if self.two_d else _FACE_AXIS_3D
        self._slots = _SLOTS_2D if self.two_d else _SLOTS_3D
        self._panels: Dict[int, List[List[int]]] = {}
        self.leaves: List = []
        self.nodes: List[NodePlan] = []  # internal nodes, deepest level first
        self.by_id: Dict[int, NodePlan] = {}
        self._collect(root)
        self.leaf_index = {id(leaf): i for i, leaf in enumerate(self.leaves)}
        internal = []
        self._internal(root, internal)
        internal.sort(key=lambda n: -n.depth)
        for node in internal:
            plan = self._plan_node(node)
            self.nodes.append(plan)
            self.by_id[id(node)] = plan

    # ---- face panel lists: per node and face, the depths of the leaves tiling it, in boundary order
    def _collect(self, node) -> List[List[int]]:
        if not node.children:
            self.leaves.append(node)
            lists = [[node.depth] for _ in range(self.n_faces)]
        else:
            kids = [self._collect(c) for c in node.children]
            lists = [[d for c in self._face_children[f] for d in kids[c][f]] for f in range(self.n_faces)]
        self._panels[id(node)] = lists
        return lists

    def _internal(self, node, out):
        if node.children:
            out.append(node)
            for c in node.children:
                self._internal(c, out)

    def n_points(self, node) -> int:
        return self.npp * sum(len(lst) for lst in self._panels[id(node)])

    def face_sizes(self, node) -> List[int]:
        return [self.npp * len(lst) for lst in self._panels[id(node)]]

    # ---- one node
    def _plan_node(self, node) -> NodePlan:
        npp, group = self.npp, self.group
        kids = node.children
        panels = [self._panels[id(c)] for c in kids]
        face_off = [np.concatenate([[0], np.cumsum([len(l) for l in p])]) for p in panels]  # native panel offsets
        # output (T') panel lists per child and face: start as plain copies, interface faces are overwritten
        out = [[[(int(face_off[c][f] + k) * npp, 1, 0) for k in range(len(panels[c][f]))] for f in range(self.n_faces)]
               for c in range(len(kids))]
        is_int = [[False] * self.n_faces for _ in kids]
        slot_faces = []
        for a, b in self._slots:
            fa, fb = _shared_faces(self._offsets, self._face_axis, a, b)
            is_int[a][fa] = is_int[b][fb] = True
            rev_b = 1 if self.two_d else 0
            la = panels[a][fa]
            lb = panels[b][fb][::-1] if rev_b else panels[b][fb]
            seg_a, seg_b = [], []
            i = j = 0
            while i < len(la) and j < len(lb):
                wa = wb = 1
                if la[i] > lb[j]:  # a's leaves are one level deeper: `group` of them face one of b's
                    wa = group
                elif la[i] < lb[j]:
                    wb = group
                seg_a.append(((int(face_off[a][fa]) + i) * npp, wa, 0))
                if rev_b:  # interface-oriented panels [j, j+wb) are native panels [n-j-wb, n-j) reversed
                    seg_b.append(((int(face_off[b][fb]) + len(lb) - j - wb) * npp, wb, 1))
                else:
                    seg_b.append(((int(face_off[b][fb]) + j) * npp, wb, 0))
                i += wa
                j += wb
            if i != len(la) or j != len(lb):
                raise ValueError("neighbouring leaves differ by more than one level: the tree is not level-restricted")
            out[a][fa], out[b][fb] = seg_a, seg_b
            slot_faces.append((a, fa, b, fb, len(seg_a)))
        children = []
        for c, kid in enumerate(kids):
            seg = np.array([s for f in range(self.n_faces) for s in out[c][f]], dtype=np.int32).reshape(-1, 3)
            off_out = np.concatenate([[0], np.cumsum([len(out[c][f]) for f in range(self.n_faces)])])
            n = self.n_points(kid)
            identity = bool(np.all(seg[:, 1] == 1) and np.all(seg[:, 2] == 0) and seg.shape[0] * npp == n
                            and np.array_equal(seg[:, 0], np.arange(seg.shape[0]) * npp))
            children.append(ChildPlan(n=n, n_out=seg.shape[0] * npp, seg=seg, identity=identity,
                                      face_off_out=[int(x) for x in off_out]))
        int_rows = []
        for a, fa, b, fb, count in slot_faces:
            for k in range(count):
                int_rows.append((a, children[a].face_off_out[fa] + k, b, children[b].face_off_out[fb] + k))
        ext_rows = []
        for f in range(self.n_faces):
            for c in self._face_children[f]:
                for k in range(len(out[c][f])):
                    ext_rows.append((c, children[c].face_off_out[f] + k))
        # non-zero blocks of B: (exterior face) x (interface face) of the same child
        slot_off = np.concatenate([[0], np.cumsum([cnt for *_, cnt in slot_faces])])
        face_slot = {}
        for s_idx, (a, fa, b, fb, _) in enumerate(slot_faces):
            face_slot[(a, fa)] = face_slot[(b, fb)] = s_idx
        ext_first, at = {}, 0
        for f in range(self.n_faces):
            for c in self._face_children[f]:
                ext_first[(c, f)] = at
                at += len(out[c][f])
        bs_rows = []
        for c in range(len(kids)):
            fo = children[c].face_off_out
            for fe in range(self.n_faces):
                if is_int[c][fe]:
                    continue
                for fi in range(self.n_faces):
                    if not is_int[c][fi]:
                        continue
                    s_idx = face_slot[(c, fi)]
                    bs_rows.append((c, fo[fe] * npp, fo[fi] * npp, (fo[fe + 1] - fo[fe]) * npp, (fo[fi + 1] - fo[fi]) * npp,
                                    int(slot_off[s_idx]) * npp, ext_first[(c, fe)] * npp))
        int_tbl = np.array(int_rows, dtype=np.int32).reshape(-1, 4)
        ext_tbl = np.array(ext_rows, dtype=np.int32).reshape(-1, 2)
        NE = ext_tbl.shape[0]
        source: Dict = {}
        for e, (c, p) in enumerate(ext_rows):
            source[(c, p)] = e
        for i, (a, pa, b, pb) in enumerate(int_rows):
            source[(a, pa)] = NE + i
            source[(b, pb)] = NE + i
        down = [(c, source[(c, P)], int(s[0]), int(s[1]), int(s[2])) for c in range(len(kids))
                for P, s in enumerate(children[c].seg)]
        return NodePlan(node=node, npp=npp, group=group, children=children, int_tbl=int_tbl, ext_tbl=ext_tbl,
                        down_tbl=np.array(down, dtype=np.int32).reshape(-1, 5),
                        bs_tbl=np.array(bs_rows, dtype=np.int32).reshape(-1, 7), n_int=int_tbl.shape[0] * npp,
                        n_ext=NE * npp, all_leaf_children=all(not k.children for k in kids))

    # ---- flatten every table into one int32 buffer (one host->device copy for the whole tree)
    def pack(self) -> np.ndarray:
        parts, at = [], 0
        for plan in self.nodes:
            if plan.all_leaf_children:
                continue
            tables = {"int": plan.int_tbl, "ext": plan.ext_tbl, "down": plan.down_tbl}
            for c, ch in enumerate(plan.children):
                tables[f"seg{c}"] = ch.seg
            for key, arr in tables.items():
                plan.off[key] = at
                parts.append(arr.reshape(-1))
                at += arr.size
        return np.concatenate(parts).astype(np.int32) if parts else np.zeros(1, dtype=np.int32)


def get_plan(pde_problem) -> TreePlan:
    """The (cached) plan of the problem's tree."""
    plan: Optional[TreePlan] = pde_problem.__dict__.get("_tree_plan")
    if plan is None:
        plan = TreePlan(pde_problem.domain.root, pde_problem.domain.q)
        pde_problem.__dict__["_tree_plan"] = plan
    return plan
