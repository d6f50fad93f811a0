"""jax_cosmo.scipy.integrate.romb restated (Romberg integration on 2**divmax + 1 points, as in jax_cosmo / the old
scipy.integrate.romberg): trapezoid sums refined by halving + Richardson extrapolation.  Test infrastructure.
Third-party arithmetic restated, not executed: jax_cosmo is absent from this image."""
import numpy as _np
from jax.numpy import _wrap


def _difftrap(function, interval, numtraps):
    if numtraps == 1:
        return 0.5 * (function(interval[0]) + function(interval[1]))
    numtosum = numtraps // 2
    h = (interval[1] - interval[0]) / numtosum
    lox = interval[0] + 0.5 * h
    points = lox + h * _np.arange(0, numtosum)
    return _np.sum(function(_wrap(points)), axis=0)


def _romberg_diff(b, c, k):
    tmp = 4.0 ** k
    return (tmp * c - b) / (tmp - 1.0)


def romb(function, a, b, args=(), divmax=6, return_error=False):
    vfunc = lambda x: function(x, *args)
    interval = [a, b]
    intrange = b - a
    ordsum = _difftrap(vfunc, interval, 1)
    result = intrange * ordsum
    state = [result] * (divmax + 1)
    err = _np.inf
    i = 0
    for i in range(1, divmax + 1):
        n = 2 ** i
        ordsum = ordsum + _difftrap(vfunc, interval, n)
        x = intrange * ordsum / n
        new_state = [x]
        for k in range(i if i < divmax else divmax):
            x = _romberg_diff(state[k], x, k + 1)
            new_state.append(x)
        while len(new_state) < divmax + 1:
            new_state.append(new_state[-1])
        err = _np.abs(state[i - 1] - new_state[i])
        state = new_state
    if return_error:
        return state[i], err
    return state[i]
