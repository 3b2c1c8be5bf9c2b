_REG = {}


def register_pytree_node(cls, flatten, unflatten):
    _REG[cls] = (flatten, unflatten)


def tree_flatten(node):
    flat = _REG[type(node)][0](node)
    return flat[0], flat[1]
