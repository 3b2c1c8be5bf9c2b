"""Discretisation-tree node types (host bookkeeping only).

Mirrors `src/jaxhps/_discretization_tree.py:6-316` minus the JAX pytree
registration: there is no tracing compiler here, nodes are plain Python objects.
"""
from __future__ import annotations

from typing import List, Tuple

__all__ = [
    "NodeData",
    "DiscretizationNode2D",
    "DiscretizationNode3D",
    "get_all_leaves",
    "get_nodes_at_level",
    "get_depth",
    "get_four_children",
    "get_eight_children",
    "add_four_children",
    "add_eight_children",
    "add_uniform_levels",
    "find_path_from_root",
    "get_ordered_lst_of_boundary_nodes",
    "get_discretization_node_area",
    "tree_equal",
    "find_node_at_corner",
    "find_nodes_along_interface_3D",
    "find_path_from_root_2D",
    "find_path_from_root_3D",
    "node_at",
    "get_all_leaves_special_ordering_3D",
    "FACE_CHILDREN_2D",
    "FACE_CHILDREN_3D",
]


class NodeData:
    """Per-node solver outputs (`_discretization_tree.py:6-32`)."""

    def __init__(self):
        self.T = None
        self.h = None
        self.S = None
        self.g_tilde = None
        self.Y = None
        self.v = None
        self.u = None
        self.g = None
        self.L_4f1 = None
        self.L_1f4 = None
        self.l2_nrm = 0.0


class DiscretizationNode2D:
    """A box of the quadtree; sides indexed 0..3 = S, E, N, W
    (`_discretization_tree.py:77-119`)."""

    def __init__(self, xmin, xmax, ymin, ymax, depth: int = 0, children: Tuple = ()):
        self.xmin = xmin
        self.xmax = xmax
        self.ymin = ymin
        self.ymax = ymax
        self.depth = depth
        self.data = NodeData()
        self.n_0 = self.n_1 = self.n_2 = self.n_3 = None
        self.children = children

    def __repr__(self):
        return "DiscretizationNode2D(xmin={}, xmax={}, ymin={}, ymax={}, depth={})".format(
            self.xmin, self.xmax, self.ymin, self.ymax, self.depth
        )


class DiscretizationNode3D:
    """A box of the octree; faces indexed 0..5 = x-, x+, y-, y+, z-, z+
    (`_discretization_tree.py:164-215`)."""

    def __init__(self, xmin, xmax, ymin, ymax, zmin, zmax, depth: int = 0, children: Tuple = ()):
        self.xmin = xmin
        self.xmax = xmax
        self.ymin = ymin
        self.ymax = ymax
        self.zmin = zmin
        self.zmax = zmax
        self.depth = depth
        self.data = NodeData()
        self.n_0 = self.n_1 = self.n_2 = self.n_3 = self.n_4 = self.n_5 = None
        self.children = children

    def __repr__(self):
        return (
            "DiscretizationNode3D(xmin={}, xmax={}, ymin={}, ymax={}, zmin={}, zmax={}, depth={})"
        ).format(self.xmin, self.xmax, self.ymin, self.ymax, self.zmin, self.zmax, self.depth)


def get_all_leaves(node) -> List:
    """Leaves in depth-first sibling order (`_discretization_tree.py:282-295`)."""
    if not node.children:
        return [node]
    out = []
    for child in node.children:
        out.extend(get_all_leaves(child))
    return out


def get_nodes_at_level(node, level: int) -> List:
    if node.depth == level:
        return [node]
    out = []
    for child in node.children:
        out.extend(get_nodes_at_level(child, level))
    return out


def get_depth(node) -> int:
    if not node.children:
        return node.depth
    return max(get_depth(c) for c in node.children)


# --------------------------------------------------------------------------- adaptive trees
#
# Children orders (reference `_discretization_tree_operations_2D.py:8-50`, `..._3D.py:102-209`):
#   2D: a SW, b SE, c NE, d NW;   3D: a(x-,y-,z+) b(x+,y-,z+) c(x+,y+,z+) d(x-,y+,z+) e..h the same at z-.
# _OFFSETS[c] = which half (0 low / 1 high) child c occupies along each axis.
_OFFSETS_2D = ((0, 0), (1, 0), (1, 1), (0, 1))
_OFFSETS_3D = ((0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1), (0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0))
# Face f of a node is tiled by face f of these children, in this order: 2D sides S,E,N,W walked
# counter-clockwise; 3D faces x-,x+,y-,y+,z-,z+ in the quad order SW,SE,NE,NW of the face's two free
# coordinates (`_grid_creation_3D.py:376-417`, `merge/_uniform_3D_DtN.py:507-541`).
FACE_CHILDREN_2D = ((0, 1), (1, 2), (2, 3), (3, 0))
FACE_CHILDREN_3D = ((4, 7, 3, 0), (5, 6, 2, 1), (4, 5, 1, 0), (7, 6, 2, 3), (4, 5, 6, 7), (0, 1, 2, 3))
# (axis, side) of each face: side 0 = the low end of the axis
_FACE_AXIS_2D = ((1, 0), (0, 1), (1, 1), (0, 0))
_FACE_AXIS_3D = ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1))


def _is_2D(node) -> bool:
    return isinstance(node, DiscretizationNode2D)


def _bounds(node):
    if _is_2D(node):
        return ((node.xmin, node.xmax), (node.ymin, node.ymax))
    return ((node.xmin, node.xmax), (node.ymin, node.ymax), (node.zmin, node.zmax))


def _make_children(parent) -> Tuple:
    two_d = _is_2D(parent)
    b = _bounds(parent)
    mid = [(lo + hi) / 2 for lo, hi in b]
    out = []
    for off in _OFFSETS_2D if two_d else _OFFSETS_3D:
        lims = []
        for ax, o in enumerate(off):
            lims += [b[ax][0], mid[ax]] if o == 0 else [mid[ax], b[ax][1]]
        out.append((DiscretizationNode2D if two_d else DiscretizationNode3D)(*lims, depth=parent.depth + 1))
    return tuple(out)


def get_four_children(parent: DiscretizationNode2D) -> Tuple:
    """The four quadrants SW, SE, NE, NW of ``parent`` (`_discretization_tree_operations_2D.py:8-50`)."""
    return _make_children(parent)


def get_eight_children(parent: DiscretizationNode3D) -> Tuple:
    """The eight octants a..h of ``parent`` (`_discretization_tree_operations_3D.py:102-209`)."""
    return _make_children(parent)


def find_path_from_root(root, node) -> List:
    """Nodes from ``root`` down to the parent of ``node`` (which must be a strict descendant), found
    by comparing ``node``'s low corner with the midpoints on the way
    (`_discretization_tree_operations_2D.py:108-146`, `..._3D.py:212-272`)."""
    if not root.children:
        raise ValueError("Specified root has no children.")
    two_d = _is_2D(root)
    offs = _OFFSETS_2D if two_d else _OFFSETS_3D
    path, cur = [], root
    while True:
        path.append(cur)
        if any(c is node for c in cur.children):
            return path
        if not cur.children:
            raise ValueError("node is not a descendant of root")
        want = tuple(int(lo >= (blo + bhi) / 2) for (lo, _), (blo, bhi) in zip(_bounds(node), _bounds(cur)))
        cur = cur.children[offs.index(want)]


def _add_children(add_to, root, q) -> None:
    if len(add_to.children):
        return  # never re-split (keeps the counts below consistent)
    add_to.children = _make_children(add_to)
    if q is None:
        return
    two_d = _is_2D(add_to)
    n_faces, per_child = (4, q) if two_d else (6, q * q)
    grow = per_child if two_d else 3 * per_child  # a leaf face becomes 2 (4) panels
    for child in add_to.children:
        for f in range(n_faces):
            setattr(child, f"n_{f}", per_child)
    for f in range(n_faces):
        setattr(add_to, f"n_{f}", per_child + grow)
    if root is None or add_to is root:
        return
    face_axis = _FACE_AXIS_2D if two_d else _FACE_AXIS_3D
    mine = _bounds(add_to)
    for anc in find_path_from_root(root, add_to):
        theirs = _bounds(anc)
        for f, (ax, side) in enumerate(face_axis):
            if theirs[ax][side] == mine[ax][side]:
                setattr(anc, f"n_{f}", getattr(anc, f"n_{f}") + grow)


def add_four_children(add_to: DiscretizationNode2D, root: DiscretizationNode2D = None, q: int = None) -> None:
    """Split a leaf into four.  With ``root`` and ``q`` the per-side Gauss-point counts ``n_0..n_3``
    of the node, its children and every ancestor sharing a side are kept up to date
    (`_discretization_tree_operations_2D.py:53-105`)."""
    _add_children(add_to, root, q)


def add_eight_children(add_to: DiscretizationNode3D, root: DiscretizationNode3D = None, q: int = None) -> None:
    """Split a leaf into eight; ``n_0..n_5`` bookkeeping as in the 2D case
    (`_discretization_tree_operations_3D.py:25-99`)."""
    _add_children(add_to, root, q)


def add_uniform_levels(root, l: int, q: int = None) -> None:
    """Refine every leaf ``l`` times — the loop the reference's adaptive tests write by hand
    (`tests/test_down_pass/test_down_pass_adaptive_3D_DtN.py:41-45`)."""
    for _ in range(l):
        for leaf in get_all_leaves(root):
            _add_children(leaf, root, q)


def get_ordered_lst_of_boundary_nodes(root) -> Tuple[List, ...]:
    """Per side (2D: S,E,N,W) or face (3D: x-,x+,y-,y+,z-,z+), the leaves touching it in the order
    their Gauss panels appear in every boundary vector
    (`_discretization_tree_operations_2D.py:149-235`, `_grid_creation_3D.py:376-417`)."""
    table = FACE_CHILDREN_2D if _is_2D(root) else FACE_CHILDREN_3D

    def walk(node, f, out):
        if not node.children:
            out.append(node)
        else:
            for c in table[f]:
                walk(node.children[c], f, out)
        return out

    return tuple(walk(root, f, []) for f in range(len(table)))


def get_discretization_node_area(node) -> float:
    """Area (2D) / volume (3D) of the box (`_discretization_tree.py:265-279`)."""
    out = 1.0
    for lo, hi in _bounds(node):
        out *= hi - lo
    return out


def tree_equal(node_a, node_b) -> bool:
    """Same box, same depth and recursively equal children (the reference compares the flattened
    pytrees, `_discretization_tree_operations_2D.py:350-358`)."""
    if type(node_a) is not type(node_b) or _bounds(node_a) != _bounds(node_b) or node_a.depth != node_b.depth:
        return False
    if len(node_a.children) != len(node_b.children):
        return False
    return all(tree_equal(a, b) for a, b in zip(node_a.children, node_b.children))


def find_node_at_corner(root: DiscretizationNode2D, xmin=None, xmax=None, ymin=None, ymax=None):
    """The LEAF of a quadtree having the given coordinates among its bounds (any subset of the four may be
    given), e.g. ``xmin=root.xmin, ymin=root.ymin`` is the SW corner leaf
    (`_discretization_tree_operations_2D.py:256-332`)."""
    want = {"xmin": xmin, "xmax": xmax, "ymin": ymin, "ymax": ymax}
    want = {k: v for k, v in want.items() if v is not None}
    hits = [leaf for leaf in get_all_leaves(root) if all(getattr(leaf, k) == v for k, v in want.items())]
    if not hits:
        raise ValueError(f"no leaf with {want}")
    return hits[0]


def find_nodes_along_interface_3D(root: DiscretizationNode3D, xval=None, yval=None, zval=None):
    """Leaves touching the plane ``x = xval`` (or y / z) from the negative and from the positive side
    (`_discretization_tree_operations_3D.py:275-336`)."""
    given = [(k, v) for k, v in (("x", xval), ("y", yval), ("z", zval)) if v is not None]
    if len(given) != 1:
        raise ValueError(f"Only one of xval, yval, or zval can be specified. Input args: {xval}, {yval}, {zval}")
    ax, val = given[0]
    leaves = get_all_leaves(root)
    neg = [leaf for leaf in leaves if getattr(leaf, ax + "max") == val]
    pos = [leaf for leaf in leaves if getattr(leaf, ax + "min") == val]
    return neg, pos


def find_path_from_root_2D(root, node):
    """Reference name of :func:`find_path_from_root` (`_discretization_tree_operations_2D.py:108-146`)."""
    return find_path_from_root(root, node)


def find_path_from_root_3D(root, node):
    """Reference name of :func:`find_path_from_root` (`_discretization_tree_operations_3D.py:212-272`)."""
    return find_path_from_root(root, node)


def node_at(node, xmin=None, xmax=None, ymin=None, ymax=None) -> bool:
    """Whether ``node`` has the given coordinates among its bounds; ``None`` entries are ignored
    (`_discretization_tree_operations_2D.py:335-347`)."""
    want = {"xmin": xmin, "xmax": xmax, "ymin": ymin, "ymax": ymax}
    return all(getattr(node, k) == v for k, v in want.items() if v is not None)


def get_all_leaves_special_ordering_3D(node, child_traversal_order=None) -> Tuple:
    """Leaves in depth-first order with the children visited in ``child_traversal_order``
    (`_discretization_tree_operations_3D.py:7-22`)."""
    order = range(8) if child_traversal_order is None else [int(c) for c in child_traversal_order]
    if not node.children:
        return (node,)
    out = ()
    for c in order:
        out += get_all_leaves_special_ordering_3D(node.children[c], child_traversal_order)
    return out
