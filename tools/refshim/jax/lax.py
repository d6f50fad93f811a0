"""jax.lax control flow as Python control flow (decisions on real parts).  Test infrastructure."""
import numpy as _np
from .numpy import _wrap, _re


def _truth(p):
    return bool(_np.all(_re(p)))


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if _truth(pred) else false_fun(*operands)


def switch(index, branches, *operands):
    i = int(_np.clip(int(_re(index)), 0, len(branches) - 1))
    return branches[i](*operands)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while _truth(cond_fun(val)):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs=None, length=None, reverse=False):
    from . import _stack_tree, _slice_tree, _tree_len
    n = _tree_len(xs) if xs is not None else int(length)
    idx = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in idx:
        carry, y = f(carry, _slice_tree(xs, i) if xs is not None else None)
        ys[i] = y
    if n == 0 or ys[0] is None:
        return carry, None
    return carry, _stack_tree(ys)


def stop_gradient(x):
    # a complex-step "tangent" lives in the imaginary part
    if isinstance(x, (tuple, list)):
        return type(x)(stop_gradient(v) for v in x)
    return _wrap(_np.real(x)) if _np.iscomplexobj(x) else x


def sign(x):
    return _wrap(_np.sign(_re(x)))
