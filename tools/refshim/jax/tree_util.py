"""Minimal jax.tree_util.  Test infrastructure."""
import functools


def register_pytree_node_class(cls):
    return cls


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if tree is None:
        return None
    return f(tree, *rest)


def tree_leaves(tree):
    if isinstance(tree, dict):
        return [l for v in tree.values() for l in tree_leaves(v)]
    if isinstance(tree, (tuple, list)):
        return [l for v in tree for l in tree_leaves(v)]
    return [] if tree is None else [tree]


Partial = functools.partial
