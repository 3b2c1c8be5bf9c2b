"""Known-answer tests of the adaptive host layer, restating the reference's own host tests on this
package's API (`tests/test_discretization_tree_operations_{2D,3D}.py`, `tests/test_adaptive_discretization_{2D,3D}.py`,
`tests/test_domain.py` of the reference)."""
import numpy as np
import pytest

import jaxhps_b200 as hps
from jaxhps_b200._adaptive_discretization import get_squared_l2_norm_single_panel, get_squared_l2_norm_single_voxel
from jaxhps_b200._grid import compute_interior_Chebyshev_points_adaptive_2D, compute_interior_Chebyshev_points_adaptive_3D
from jaxhps_b200._tree import (
    add_eight_children,
    add_four_children,
    add_uniform_levels,
    find_path_from_root,
    get_all_leaves,
    get_eight_children,
    get_four_children,
    get_ordered_lst_of_boundary_nodes,
)


def area(n):
    v = (n.xmax - n.xmin) * (n.ymax - n.ymin)
    return v * (n.zmax - n.zmin) if hasattr(n, "zmin") else v


def test_children_geometry_and_order():
    kids = get_four_children(hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0))
    assert [(k.xmin, k.ymin) for k in kids] == [(0.0, 0.0), (0.5, 0.0), (0.5, 0.5), (0.0, 0.5)]  # SW SE NE NW
    assert all(area(k) == 0.25 and k.depth == 1 and k.children == () for k in kids)
    kids = get_eight_children(hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0))
    assert [(k.xmin, k.ymin, k.zmin) for k in kids] == [(0, 0, .5), (.5, 0, .5), (.5, .5, .5), (0, .5, .5),
                                                       (0, 0, 0), (.5, 0, 0), (.5, .5, 0), (0, .5, 0)]
    assert all(area(k) == 1 / 8 for k in kids)


@pytest.mark.parametrize("l", [1, 2, 3])
def test_uniform_refinement_counts_2D(l):
    """reference test_discretization_tree_operations_2D.py:137-183"""
    q = 3
    root = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    add_uniform_levels(root, l, q)
    assert [getattr(root, f"n_{f}") for f in range(4)] == [2**l * q] * 4
    for c in root.children:
        assert [getattr(c, f"n_{f}") for f in range(4)] == [2 ** (l - 1) * q] * 4
    assert len(get_all_leaves(root)) == 4**l


def test_nonuniform_counts_2D_and_3D():
    """reference 2D test_2 (`:103-135`) and 3D test_2 (`:115-140`)."""
    q = 3
    root = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    add_four_children(root, root=root, q=q)
    for c in root.children:
        add_four_children(c, root=root, q=q)
    add_four_children(root.children[0].children[0], root=root, q=q)
    assert (root.n_0, root.n_1, root.n_2, root.n_3) == (5 * q, 4 * q, 4 * q, 5 * q)
    q = 4
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    add_eight_children(root, root=root, q=q)
    add_eight_children(root.children[0], root=root, q=q)
    add_eight_children(root.children[0].children[0], root=root, q=q)
    assert (root.n_1, root.n_3, root.n_4) == (4 * q * q,) * 3
    assert (root.n_0, root.n_2, root.n_5) == (10 * q * q,) * 3
    assert len(get_all_leaves(root)) == 22
    # uniform 3D, two levels: 16 panels per face (reference 3D test_1 with l=2 gives 64 for l=3)
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    add_uniform_levels(root, 3, q)
    assert [getattr(root, f"n_{f}") for f in range(6)] == [64 * q * q] * 6


def test_path_from_root_and_boundary_walks():
    root = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    add_uniform_levels(root, 2)
    node = root.children[2].children[1]
    path = find_path_from_root(root, node)
    assert len(path) == 2 and path[0] is root and path[1] is root.children[2]
    with pytest.raises(ValueError):
        find_path_from_root(hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0), node)
    # counter-clockwise walk from the SW corner on a non-uniform quadtree
    root = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    add_four_children(root)
    add_four_children(root.children[1])
    S, E, N, W = get_ordered_lst_of_boundary_nodes(root)
    assert [(n.xmin, n.xmax) for n in S] == [(0.0, 0.5), (0.5, 0.75), (0.75, 1.0)]
    assert [(n.ymin, n.ymax) for n in E] == [(0.0, 0.25), (0.25, 0.5), (0.5, 1.0)]
    assert [(n.xmin, n.xmax) for n in N] == [(0.5, 1.0), (0.0, 0.5)]
    assert [(n.ymin, n.ymax) for n in W] == [(0.5, 1.0), (0.0, 0.5)]
    # 3D: every face of the reference's test tree lists its leaves once
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    add_eight_children(root)
    add_eight_children(root.children[6])
    faces = get_ordered_lst_of_boundary_nodes(root)
    assert [len(f) for f in faces] == [4, 7, 4, 7, 7, 4]
    assert all(leaf.xmax == 1.0 for leaf in faces[1]) and all(leaf.zmin == 0.0 for leaf in faces[4])


def test_squared_l2_norms():
    """reference test_adaptive_discretization_3D.py:84-135 and the 2D analogues."""
    root = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    b = [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    assert np.isclose(get_squared_l2_norm_single_voxel(3 * np.ones(4**3), b, 4), 9.0)
    pts = compute_interior_Chebyshev_points_adaptive_3D(root, 4)[0]
    assert np.isclose(get_squared_l2_norm_single_voxel(np.sqrt(pts[:, 0] + pts[:, 1] + pts[:, 2]), b, 4), 1.5)
    pts = compute_interior_Chebyshev_points_adaptive_3D(root, 6)[0]
    assert np.isclose(get_squared_l2_norm_single_voxel(pts[:, 0] ** 2 + pts[:, 1], b, 6), 13 / 15)
    root2 = hps.DiscretizationNode2D(0.0, 2.0, -1.0, 1.0)
    pts = compute_interior_Chebyshev_points_adaptive_2D(root2, 6)[0]
    # ||x + y||^2 over [0,2]x[-1,1] = int (x+y)^2 = 8/3*2 ... computed exactly: 16/3 + 0 + 4/3
    assert np.isclose(get_squared_l2_norm_single_panel(pts[:, 0] + pts[:, 1], [0.0, 2.0, -1.0, 1.0], 6), 16 / 3 + 4 / 3)


def test_adaptive_mesh_resolves_function_and_is_level_restricted():
    """The generated mesh meets its tolerance and neighbouring leaves differ by at most one level."""
    def f(x):
        return np.exp(-60 * ((x[..., 0] - 0.2) ** 2 + (x[..., 1] - 0.3) ** 2))

    p, q, tol = 8, 6, 1e-4
    root = hps.DiscretizationNode2D(-1.0, 1.0, -1.0, 1.0)
    dom = hps.Domain.from_adaptive_discretization(p=p, q=q, root=root, f=f, tol=tol)
    leaves = get_all_leaves(root)
    assert dom.n_leaves == len(leaves) > 4 and not dom.bool_uniform
    # interpolation from the leaf grids reproduces f to about the tolerance
    xs = np.linspace(-0.99, 0.99, 41)
    vals, pts = dom.interp_from_interior_points(f(dom.interior_points), xs, xs)
    assert np.abs(vals - f(pts)).max() < 20 * tol
    # level restriction: any two leaves sharing part of an edge differ by at most one level
    for a in leaves:
        for b in leaves:
            touch_x = (a.xmax == b.xmin) and (min(a.ymax, b.ymax) - max(a.ymin, b.ymin) > 0)
            touch_y = (a.ymax == b.ymin) and (min(a.xmax, b.xmax) - max(a.xmin, b.xmin) > 0)
            if touch_x or touch_y:
                assert abs(a.depth - b.depth) <= 1
    # adaptive problems need a source at build time (reference `_pdeproblem.py:77-82`)
    with pytest.raises(ValueError):
        hps.PDEProblem(dom, source=None)
    # a tree that is NOT level-restricted is refused by the merge-plan compiler
    from jaxhps_b200._adaptive_plan import TreePlan

    bad = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    add_four_children(bad, root=bad, q=q)
    add_four_children(bad.children[0], root=bad, q=q)
    add_four_children(bad.children[0].children[1], root=bad, q=q)  # depth-3 leaves next to the depth-1 leaf SE
    with pytest.raises(ValueError):
        TreePlan(bad, q)


def test_uniform_tree_given_as_adaptive_domain_matches_uniform_domain():
    """reference test_domain.py: an `L=None` Domain over a uniformly refined tree has the uniform point clouds."""
    for dim in (2, 3):
        if dim == 2:
            root, root_u = hps.DiscretizationNode2D(-1.0, 1.0, 0.0, 2.0), hps.DiscretizationNode2D(-1.0, 1.0, 0.0, 2.0)
        else:
            root, root_u = (hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, -1.0, 0.0) for _ in range(2))
        add_uniform_levels(root, 2, 4)
        a, u = hps.Domain(6, 4, root), hps.Domain(6, 4, root_u, L=2)
        assert np.array_equal(a.interior_points, u.interior_points)
        np.testing.assert_allclose(a.boundary_points, u.boundary_points, rtol=0, atol=1e-15)
        lst = a.get_adaptive_boundary_data_lst(lambda x: x[..., 0])
        assert sum(len(g) for g in lst) == u.boundary_points.shape[0]


def test_tree_queries():
    """reference test_discretization_tree_operations_2D.py:188-250, ..._3D.py:183-238."""
    from jaxhps_b200._tree import (find_node_at_corner, find_nodes_along_interface_3D, get_discretization_node_area,
                                   tree_equal)

    root = hps.DiscretizationNode2D(0.0, 1.0, 0.0, 1.0)
    add_uniform_levels(root, 2)
    sw = find_node_at_corner(root, xmin=0.0, ymin=0.0)
    assert tree_equal(sw, root.children[0].children[0]) and get_discretization_node_area(sw) == 1 / 16
    ne = find_node_at_corner(root, xmax=1.0, ymax=1.0)
    assert ne is root.children[2].children[2] and not tree_equal(sw, ne)
    assert find_node_at_corner(root, xmin=0.25, ymin=0.0) is root.children[0].children[1]
    root3 = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    add_uniform_levels(root3, 2)
    for kw in ({"xval": 0.5}, {"yval": 0.5}, {"zval": 0.5}):
        neg, pos = find_nodes_along_interface_3D(root3, **kw)
        assert len(neg) == len(pos) == 16
    root3 = hps.DiscretizationNode3D(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
    add_eight_children(root3)
    add_eight_children(root3.children[0])
    neg, pos = find_nodes_along_interface_3D(root3, xval=0.5)  # child a is refined: 4 small + 3 big leaves on its side
    assert (len(neg), len(pos)) == (7, 4)
    with pytest.raises(ValueError):
        find_nodes_along_interface_3D(root3, xval=0.5, yval=0.5)
    import jaxhps_b200.local_solve as ls
    import jaxhps_b200.merge as mg

    for name in ("local_solve_stage_adaptive_3D_DtN", "nosource_local_solve_stage_uniform_2D_ItI"):
        assert callable(getattr(ls, name))
    for name in ("merge_stage_adaptive_2D_DtN", "nosource_merge_stage_uniform_2D_DtN"):
        assert callable(getattr(mg, name))
