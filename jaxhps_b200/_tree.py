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
