"""jax.flatten_util.ravel_pytree for arrays / flat containers.  Test infrastructure."""
import numpy as _np
from .numpy import _wrap
from .tree_util import tree_leaves


def ravel_pytree(tree):
    leaves = tree_leaves(tree)
    flat = _np.concatenate([_np.ravel(_np.asarray(l)) for l in leaves]) if leaves else _np.zeros(0)
    return _wrap(flat), None
