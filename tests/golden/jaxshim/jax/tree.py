"""`jax.tree.map` for the containers the reference passes it (lists/tuples/dicts of nodes)."""


def map(f, tree, *rest, is_leaf=None):  # noqa: A001
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    if isinstance(tree, (list, tuple)):
        out = [map(f, t, *(r[i] for r in rest), is_leaf=is_leaf) for i, t in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    if isinstance(tree, dict):
        return {k: map(f, v, *(r[k] for r in rest), is_leaf=is_leaf) for k, v in tree.items()}
    if tree is None:
        return None
    return f(tree, *rest)
