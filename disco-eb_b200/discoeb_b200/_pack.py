"""Packs the reference's ``param`` dict into the flat float64 blocks of the C-ABI
(``include/discoeb_b200.h``: ``scalars[DEB_NSCAL]`` and ``tables[deb_table_len]`` per cosmology).

The reference keeps the cosmology in a plain dict of scalars plus ``spline_interpolation``
pytrees (``/root/reference/src/discoeb/spline_interpolation.py:111-121``: ``_x_``, ``_y_``,
``_S_full_``).  Any object exposing either those attributes or ``x``, ``y``, ``S`` is accepted, so
dicts produced by the real ``evolve_background`` (JAX arrays) and by the test-suite's table
producer both work.
"""
from __future__ import annotations

import numpy as np

NSCAL = 24
SCALAR_KEYS = ("Omegam", "Omegab", "OmegaDE", "Omegak", "grhom", "grhog", "grhor", "Neff", "Nmnu", "amnu",
               "w_DE_0", "w_DE_a", "cs2_DE", "YHe", "H0", "taumin", "A_s", "n_s", "k_p")
SCALAR_DEFAULTS = {"A_s": 1.0, "n_s": 1.0, "k_p": 0.05, "Omegak": 0.0}
SPLINE_KEYS = ("cs2a_of_loga_spline", "xe_of_loga_spline", "logrhonu_of_loga_spline", "logpnu_of_loga_spline",
               "a_of_tau_spline", "xe_of_tau_spline", "tau_of_a_spline")


def _spline_arrays(sp):
    if hasattr(sp, "_x_"):
        x, y, S = sp._x_, sp._y_, sp._S_full_
    else:
        x, y, S = sp.x, sp.y, sp.S
    return (np.ascontiguousarray(np.asarray(x, dtype=np.float64)),
            np.ascontiguousarray(np.asarray(y, dtype=np.float64)),
            np.ascontiguousarray(np.asarray(S, dtype=np.float64)))


def pack_param(param):
    """-> (scalars[NSCAL], tables[3*(5*nth+2*nnu)], nth, nnu) for one cosmology."""
    scal = np.zeros(NSCAL, dtype=np.float64)
    for i, key in enumerate(SCALAR_KEYS):
        if key in param:
            scal[i] = float(param[key])
        elif key in SCALAR_DEFAULTS:
            scal[i] = SCALAR_DEFAULTS[key]
        else:
            raise KeyError(f"param['{key}'] is required by evolve_perturbations (run evolve_background first)")
    parts = []
    sizes = []
    for key in SPLINE_KEYS:
        if key not in param:
            raise KeyError(f"param['{key}'] is required by evolve_perturbations (run evolve_background first)")
        x, y, S = _spline_arrays(param[key])
        if not (x.shape == y.shape == S.shape and x.ndim == 1 and x.shape[0] >= 2):
            raise ValueError(f"{key}: x, y, S must be 1-d arrays of equal length >= 2")
        parts += [x, y, S]
        sizes.append(x.shape[0])
    nth, nnu = sizes[0], sizes[2]
    if not (sizes[1] == nth and sizes[4] == nth and sizes[5] == nth and sizes[6] == nth and sizes[3] == nnu):
        raise ValueError("thermo splines must share one knot count and the two neutrino splines another")
    if not np.array_equal(parts[0], parts[3]):
        raise ValueError("cs2a_of_loga_spline and xe_of_loga_spline must share their knots")
    if not np.array_equal(parts[6], parts[9]):
        raise ValueError("logrhonu_of_loga_spline and logpnu_of_loga_spline must share their knots")
    return scal, np.concatenate(parts), nth, nnu


def pack_params(params):
    """Stack several cosmologies: -> (scalars[nc, NSCAL], tables[nc, tl], nth, nnu)."""
    packed = [pack_param(p) for p in params]
    nth, nnu = packed[0][2], packed[0][3]
    for _, _, a, b in packed:
        if (a, b) != (nth, nnu):
            raise ValueError("all cosmologies of a batch must use the same table sizes")
    return (np.ascontiguousarray(np.stack([p[0] for p in packed])),
            np.ascontiguousarray(np.stack([p[1] for p in packed])), nth, nnu)


def pack_tangent(param, dparam):
    """Tangent seed of one cosmology along one direction -> (d_scalars[NSCAL], d_tables[tl]) in the layout of
    ``pack_param``.  ``dparam`` maps any subset of the scalar keys to floats and of the spline keys to objects
    carrying the tangents of ``x, y, S`` (what ``jax.jvp`` holds for the ``param`` pytree); missing keys have
    zero tangent."""
    dscal = np.zeros(NSCAL, dtype=np.float64)
    for i, key in enumerate(SCALAR_KEYS):
        if key in dparam:
            dscal[i] = float(dparam[key])
    parts = []
    for key in SPLINE_KEYS:
        x, y, S = _spline_arrays(param[key])
        if key in dparam:
            dx, dy, dS = _spline_arrays(dparam[key])
            if not (dx.shape == x.shape and dy.shape == y.shape and dS.shape == S.shape):
                raise ValueError(f"dparam['{key}']: tangent arrays must have the shapes of the primal spline")
            parts += [dx, dy, dS]
        else:
            parts += [np.zeros_like(x), np.zeros_like(y), np.zeros_like(S)]
    return dscal, np.concatenate(parts)
