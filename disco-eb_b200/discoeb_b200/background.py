"""``evolve_background`` on the GPU: the tables ``evolve_perturbations`` reads, for one cosmology or a batch.

Mirrors the reference's ``evolve_background(*, param, thermo_module='RECFAST', num_thermo=256)``
(/root/reference/src/discoeb/background.py:191-342): takes the same ``param`` dict (Omegam, Omegab, w_DE_0, w_DE_a,
cs2_DE, Omegak, A_s, n_s, H0, Tcmb, YHe, Neff, Nmnu, mnu) and returns it with the derived scalars (grhom, grhog, grhor,
amnu, OmegaDE, Omegamnu, taumin, taumax) and the seven splines of the hot path (``*_spline`` objects carrying
``x, y, S`` and an ``evaluate`` method).  Everything is computed by ``deb_background_f64`` (csrc/deb_background.cuh), one
CTA per cosmology; there is no CPU path.  Not produced: optical depth / visibility (background.py:300-342), which
``evolve_perturbations`` does not read, and the MB95 / CLASS thermodynamics modules.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from ._pack import SCALAR_KEYS, SPLINE_KEYS

BG_KEYS = ("Omegam", "Omegab", "Omegak", "w_DE_0", "w_DE_a", "cs2_DE", "H0", "Tcmb", "YHe", "Neff", "Nmnu", "mnu", "A_s", "n_s", "k_p")
BG_DEFAULTS = {"Omegak": 0.0, "A_s": 1.0, "n_s": 1.0, "k_p": 0.05, "cs2_DE": 1.0, "w_DE_0": -1.0, "w_DE_a": 0.0}
NBGIN = 16
NNU = 512


class TableSpline:
    """Natural cubic spline given by its knots and second derivatives (the reference's ``spline_interpolation``
    pytree, spline_interpolation.py:111-153); ``_x_/_y_/_S_full_`` aliases keep reference-style consumers working."""

    def __init__(self, x, y, S):
        self.x, self.y, self.S = np.ascontiguousarray(x), np.ascontiguousarray(y), np.ascontiguousarray(S)

    _x_ = property(lambda self: self.x)
    _y_ = property(lambda self: self.y)
    _S_full_ = property(lambda self: self.S)

    def evaluate(self, xn):
        xn = np.asarray(xn, dtype=np.float64)
        n = self.x.shape[0]
        i = np.clip(np.searchsorted(self.x, xn) - 1, 0, n - 2)
        h = self.x[i + 1] - self.x[i]
        t = (xn - self.x[i]) / h
        A, B = 1 - t, t
        return A * self.y[i] + B * self.y[i + 1] + ((A ** 3 - A) * self.S[i] + (B ** 3 - B) * self.S[i + 1]) * h ** 2 / 6.0

    # the other members of the reference's spline_interpolation (spline_interpolation.py:155-260), so that a ``param``
    # returned by evolve_background serves the reference's callers of these splines too
    def _local(self, xn):
        xn = np.atleast_1d(np.asarray(xn, dtype=np.float64))
        n = self.x.shape[0]
        i = np.clip(np.searchsorted(self.x, xn) - 1, 0, n - 2)
        h = self.x[i + 1] - self.x[i]
        b = (self.y[i + 1] - self.y[i]) / h - h * (self.S[i + 1] + 2 * self.S[i]) / 6.0
        return xn, i, xn - self.x[i], b, self.S[i] / 2.0, (self.S[i + 1] - self.S[i]) / (6.0 * h)

    def derivative12(self, xn):
        xn, i, dv, b, c, d = self._local(xn)
        d1, d2 = b + 2 * c * dv + 3 * d * dv ** 2, 2 * c + 6 * d * dv
        return (d1[0], d2[0]) if xn.shape[0] == 1 else (d1, d2)

    def derivative(self, xn):
        return self.derivative12(xn)[0]

    def derivative2(self, xn):
        return self.derivative12(xn)[1]

    def integral(self, xn):
        """Integral from x[0] to xn (``integrate_from_start``, the default) or from xn to x[-1]."""
        xn, i, dv, b, c, d = self._local(xn)
        h = np.diff(self.x)
        bb = np.diff(self.y) / h - h * (self.S[1:] + 2 * self.S[:-1]) / 6.0
        full = self.y[:-1] * h + bb * h ** 2 / 2 + (self.S[:-1] / 2.0) * h ** 3 / 3 + ((self.S[1:] - self.S[:-1]) / (6.0 * h)) * h ** 4 / 4
        icum = np.concatenate([[0.0], np.cumsum(full)])
        fwd = icum[i] + self.y[i] * dv + b * dv ** 2 / 2 + c * dv ** 3 / 3 + d * dv ** 4 / 4
        res = fwd if getattr(self, "integrate_from_start", True) else icum[-1] - fwd
        return res[0] if xn.shape[0] == 1 else res


def pack_background_input(param):
    v = np.zeros(NBGIN)
    for i, key in enumerate(BG_KEYS):
        if key in param:
            v[i] = float(param[key])
        elif key in BG_DEFAULTS:
            v[i] = BG_DEFAULTS[key]
        else:
            raise KeyError(f"param['{key}'] is required by evolve_background")
    return v


def config4_draws(ncosmo, seed=0):
    """BASELINE config 4: independent uniform draws of numpy.random.default_rng(seed)."""
    rng = np.random.default_rng(seed)
    lo_hi = dict(Omegam=(0.25, 0.40), Omegab=(0.04, 0.06), h=(0.60, 0.75), n_s=(0.92, 1.00), A_s=(1.7e-9, 2.5e-9), mnu=(0.0, 0.3),
                 w_DE_0=(-0.95, -0.75), w_DE_a=(0.0, 0.3))
    cols = {k: rng.uniform(lo, hi, size=ncosmo) for k, (lo, hi) in lo_hi.items()}
    out = []
    for i in range(ncosmo):
        d = {k: float(v[i]) for k, v in cols.items()}
        d["H0"] = 100.0 * d.pop("h")
        out.append(d)
    return out


def background_tables(params, *, num_thermo: int = 256, device: int = 0, lib=None, return_info: bool = False):
    """Batch entry: list of ``param`` dicts -> (scalars[nc, 24], tables[nc, 3 (5 nth + 2 nnu)]) in the packed layout of
    ``deb_evolve_f64`` (include/discoeb_b200.h), ready for ``evolve_perturbations`` without a dict round trip."""
    lib = lib or _cabi.default_library()
    bg_in = np.ascontiguousarray(np.stack([pack_background_input(p) for p in params]))
    scal, tab, ms = lib.background_host(bg_in, num_thermo, device=device)
    return (scal, tab, dict(kernel_ms=ms)) if return_info else (scal, tab)


EXTRA_ROWS = ("xeHI", "xeHeI", "xeHeII", "cs2", "Tm", "xeprime_recf", "xeprime", "opac", "optical_depth", "gvis", "gvisprime", "gvispprime")


def unpack_extras(out, ext, nth, nnu=NNU):
    """The rest of the reference's ``param`` (background.py:238-251, 167, 306-342) from the extras block of
    ``deb_background_ex_f64`` (row order: include/discoeb_b200.h)."""
    for j, key in enumerate(EXTRA_ROWS):
        out[key] = ext[j * nth:(j + 1) * nth]
    r = len(EXTRA_ROWS)
    tau, aexp = out["tau"], out["aexp"]
    out["cs2a_of_tau_spline"] = TableSpline(tau, aexp * out["cs2"], ext[r * nth:(r + 1) * nth])
    out["tempba_of_tau_spline"] = TableSpline(tau, aexp * out["Tm"], ext[(r + 1) * nth:(r + 2) * nth])
    o = (r + 2) * nth
    out["logppseudonu_of_loga_spline"] = TableSpline(out["logrhonu_of_loga_spline"].x, ext[o:o + nnu], ext[o + nnu:o + 2 * nnu])
    out["a"] = np.exp(out["logrhonu_of_loga_spline"].x)           # the knots of the neutrino tables (background.py:158-160)
    out["adotrad"] = float(np.sqrt((out["grhog"] + out["grhor"] * (out["Neff"] + out["Nmnu"])) / 3.0))
    out["fHe"] = out["YHe"] / (3.97146570884 * (1.0 - out["YHe"]))      # thermodynamics_recfast.py:41, 457
    return out


def unpack_param(param, scal, tab, nth, nnu=NNU):
    out = dict(param)
    for i, key in enumerate(SCALAR_KEYS):
        out[key] = float(scal[i])
    out["taumax"], out["Omegamnu"] = float(scal[19]), float(scal[20])
    off = 0
    for j, key in enumerate(SPLINE_KEYS):
        n = nnu if j in (2, 3) else nth
        out[key] = TableSpline(tab[off:off + n], tab[off + n:off + 2 * n], tab[off + 2 * n:off + 3 * n])
        off += 3 * n
    out["aexp"], out["tau"] = out["tau_of_a_spline"].x, out["tau_of_a_spline"].y
    out["xe"] = out["xe_of_tau_spline"].y
    out["amin"], out["amax"] = 1e-9, 1.01
    return out


def evolve_background(*, param, thermo_module: str = "RECFAST", num_thermo: int = 256, rtol: float = 1e-5, atol: float = 1e-7,
                      order: int = 5, class_thermo=None, device: int = 0, lib=None):
    """background.py:191-342 for ``thermo_module='RECFAST'`` (the reference's default and the only module the
    hot path's tests use).  ``rtol/atol/order`` are accepted and, as in the reference, unused by the RECFAST branch."""
    if thermo_module != "RECFAST":
        raise NotImplementedError("discoeb_b200.evolve_background implements thermo_module='RECFAST' only")
    lib = lib or _cabi.default_library()
    bg_in = np.ascontiguousarray(pack_background_input(param)[None])
    scal, tab, _, ext = lib.background_host(bg_in, num_thermo, device=device, extras=True)
    return unpack_extras(unpack_param(param, scal[0], tab[0], num_thermo), ext[0], num_thermo)
