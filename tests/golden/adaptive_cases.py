"""Recipes of the adaptive golden cases, shared by the generator (which runs them on the reference)
and by the tests (which run them on this repo).  ``api`` is the package to build with: it must
provide DiscretizationNode2D/3D, Domain, add_four_children / add_eight_children."""
import numpy as np

ADAPTIVE_CASES = {
    # name: dim, p, q, how the tree is made, seed of the coefficient fields
    "adapt2d_p8q6": dict(dim=2, p=8, q=6, how="generate", tol=1e-3, l2=False, seed=41),
    "adapt2d_p6q4_l2": dict(dim=2, p=6, q=4, how="generate", tol=3e-2, l2=True, seed=42),
    "adapt2d_p6q4_manual": dict(dim=2, p=6, q=4, how="manual", path=[[1], [1, 1], [3]], seed=45),
    "adapt3d_p4q2_manual": dict(dim=3, p=4, q=2, how="manual", path=[[2], [2, 2]], seed=43),
    "adapt3d_p6q4": dict(dim=3, p=6, q=4, how="generate", tol=1e-2, l2=False, seed=44),
    "adapt3d_p4q2_l2": dict(dim=3, p=4, q=2, how="generate", tol=5e-2, l2=True, seed=46),
    # random level-restricted trees (split sequences recorded from a seeded random refinement): 49 and 43 leaves
    "adapt2d_p6q4_random": dict(dim=2, p=6, q=4, how="manual", seed=47,
                                path=[[0], [1], [2], [3], [0, 0], [0, 2], [0, 3], [1, 0], [1, 1], [1, 2], [2, 0], [3, 1],
                                      [0, 0, 0], [1, 1, 2], [1, 2, 1]]),
    "adapt3d_p4q2_random": dict(dim=3, p=4, q=2, how="manual", seed=48, path=[[3], [4], [6], [7], [7, 6]]),
}


def bump(x, xp=np):
    """The function the trees are refined on ([..., d] -> [...])."""
    if x.shape[-1] == 2:
        return xp.exp(-40 * ((x[..., 0] - 0.3) ** 2 + (x[..., 1] + 0.2) ** 2))
    return xp.exp(-30 * ((x[..., 0] - 0.3) ** 2 + (x[..., 1] - 0.2) ** 2 + (x[..., 2] - 0.7) ** 2))


def boundary_fn(x, xp=np):
    if x.shape[-1] == 2:
        return xp.sin(2 * x[..., 0]) + 2 * x[..., 1] * x[..., 0]
    return xp.sin(2 * x[..., 0]) + 2 * x[..., 1] * x[..., 2]


def make_root(api, dim):
    if dim == 2:
        return api.DiscretizationNode2D(xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0)
    return api.DiscretizationNode3D(xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, zmin=0.0, zmax=1.0)


def build_domain(api, add_children, case, xp=np):
    """add_children(node, root, q) splits a leaf."""
    root = make_root(api, case["dim"])
    if case["how"] == "generate":
        return api.Domain.from_adaptive_discretization(
            p=case["p"], q=case["q"], root=root, f=lambda x: bump(x, xp), tol=case["tol"], use_l_2_norm=case["l2"])
    add_children(root, root, case["q"])
    for path in case["path"]:
        node = root
        for c in path:
            node = node.children[c]
        add_children(node, root, case["q"])
    return api.Domain(p=case["p"], q=case["q"], root=root)


def seeded_fields(case, n_leaves):
    rng = np.random.default_rng(case["seed"])
    shp = (n_leaves, case["p"] ** case["dim"])
    co = {"D_xx_coefficients": 1 + 0.1 * rng.normal(size=shp), "D_yy_coefficients": np.ones(shp),
          "D_x_coefficients": rng.normal(size=shp), "I_coefficients": rng.normal(size=shp)}
    if case["dim"] == 3:
        co["D_zz_coefficients"] = 1 + 0.1 * rng.normal(size=shp)
        co["D_yz_coefficients"] = 0.1 * rng.normal(size=shp)
    else:
        co["D_xy_coefficients"] = 0.1 * rng.normal(size=shp)
    return co, rng.normal(size=shp)


def internal_nodes(root):
    """Nodes with children, pre-order: the order per-node golden arrays are stored in."""
    out = []

    def walk(n):
        if len(n.children):
            out.append(n)
            for c in n.children:
                walk(c)

    walk(root)
    return out
