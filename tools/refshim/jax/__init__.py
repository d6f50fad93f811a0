"""jax stand-in on NumPy (see ../README.md).  Test infrastructure: runs the reference's own Python on the CPU."""
import functools as _ft
import numpy as _np

from . import numpy as numpy            # noqa: F401  (jax.numpy)
from . import lax, tree_util, flatten_util, scipy   # noqa: F401
from .numpy import ndarray as Array, _wrap

__version__ = "0.0-refshim"
CSTEP = 1e-30        # complex-step size of jacfwd / grad / jvp


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def print_environment_info():
    print("refshim: jax stand-in on NumPy", _np.__version__)


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _is_leaf(x):
    return not isinstance(x, (tuple, list, dict))


def _stack_tree(items):
    first = items[0]
    if isinstance(first, tuple):
        return tuple(_stack_tree([it[i] for it in items]) for i in range(len(first)))
    if isinstance(first, list):
        return [_stack_tree([it[i] for it in items]) for i in range(len(first))]
    if isinstance(first, dict):
        return {k: _stack_tree([it[k] for it in items]) for k in first}
    if first is None:
        return None
    return _wrap(_np.stack([_np.asarray(it) for it in items]))


def _slice_tree(x, i):
    if isinstance(x, tuple):
        return tuple(_slice_tree(v, i) for v in x)
    if isinstance(x, list):
        return [_slice_tree(v, i) for v in x]
    if isinstance(x, dict):
        return {k: _slice_tree(v, i) for k, v in x.items()}
    return x[i]


def _tree_len(x):
    if isinstance(x, (tuple, list)):
        return _tree_len(x[0])
    if isinstance(x, dict):
        return _tree_len(next(iter(x.values())))
    return len(x)


def vmap(fun, in_axes=0, out_axes=0):
    """Python loop over the mapped leading axis (in_axes entries: 0 or None)."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                assert ax == 0, "refshim.vmap maps axis 0 only"
                n = _tree_len(a)
                break
        outs = []
        for i in range(n):
            call = [(_slice_tree(a, i) if ax is not None else a) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        return _stack_tree(outs)
    return mapped


def jacfwd(fun, argnums=0):
    """Complex-step Jacobian: d fun / d x_j = Im fun(x + i h e_j) / h, exact to round-off for analytic fun."""
    def jac(*args):
        x = _np.asarray(args[argnums])
        if _np.iscomplexobj(x):
            raise TypeError("refshim.jacfwd cannot be nested inside another complex-step derivative")
        rest = list(args)
        if x.ndim == 0:
            rest[argnums] = complex(float(x), CSTEP)
            return _wrap(_np.asarray(_np.imag(fun(*rest)) / CSTEP))
        cols = []
        for j in range(x.size):
            xp = x.astype(complex).view(numpy.ndarray)
            xp[j] += 1j * CSTEP
            rest[argnums] = xp
            cols.append(_np.imag(_np.asarray(fun(*rest))) / CSTEP)
        return _wrap(_np.stack(cols, axis=-1))
    return jac


def grad(fun, argnums=0):
    def g(*args):
        x = _np.asarray(args[argnums], dtype=float)
        rest = list(args)
        if x.ndim == 0:
            rest[argnums] = complex(float(x), CSTEP)
            return float(_np.imag(fun(*rest)) / CSTEP)
        out = _np.zeros_like(x)
        for j in range(x.size):
            xp = x.astype(complex).view(numpy.ndarray)
            xp.flat[j] += 1j * CSTEP
            rest[argnums] = xp
            out.flat[j] = _np.imag(fun(*rest)) / CSTEP
        return _wrap(out)
    return g


class custom_jvp:
    """Only the primal function is ever called here."""
    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        _ft.update_wrapper(self, fun)

    def defjvp(self, jvp):
        self.jvp = jvp
        return jvp

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


def checkpoint(fun=None, **kw):
    return fun if fun is not None else (lambda f: f)
