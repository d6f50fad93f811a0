"""Spectra post-processing entry points of the reference's host API
(``/root/reference/src/discoeb/perturbations.py:1063-1224``), computed by the CUDA library
(``csrc/deb_spectra.cu``: ``deb_spectra_host_f64``; the device-pointer form ``deb_spectra_f64`` chains behind the solve on
the same stream).  Same names, keyword arguments and return values as the reference; there is no CPU path.

=========================  ============================================
``get_power``               perturbations.py:1101-1123  (also fused into the solve kernel: ``power_idx`` of the C-ABI)
``get_power_smoothed``      perturbations.py:1126-1160  (util.savgol_filter :407-444)
``power_Kaiser``            perturbations.py:1162-1199
``power_multipoles``        perturbations.py:1202-1224
``get_xi_from_P``           perturbations.py:1063-1098  (FFTlog; util.lngamma_complex_e :12-44)
=========================  ============================================
"""
from __future__ import annotations

import numpy as np

from . import _cabi

__all__ = ["get_power", "get_power_smoothed", "power_Kaiser", "power_multipoles", "get_xi_from_P"]


def get_power(*, k, y, idx: int, param):
    """``2 pi^2 A_s (k/k_p)^(n_s-1) k^-3 y[..., idx]^2`` -- one broadcast expression on whatever array shape the caller
    holds (the solve kernel's epilogue writes the same quantity when ``power_idx`` is set)."""
    k = np.asarray(k)
    y = np.asarray(y)
    return 2 * np.pi ** 2 * param["A_s"] * (k / param["k_p"]) ** (param["n_s"] - 1) * k ** (-3) * y[..., idx] ** 2


def _sg_weights(window_length: int, polyorder: int = 3):
    """Centre weights of the least-squares polynomial fit (util.py:407-444); a (polyorder+1)-unknown host solve."""
    halflen, rem = divmod(int(window_length), 2)
    pos = halflen - 0.5 if rem == 0 else halflen
    x = np.arange(-pos, window_length - pos, dtype=float)[::-1]
    A = x ** np.arange(polyorder + 1).reshape(-1, 1)
    Y = np.zeros(polyorder + 1)
    Y[0] = 1.0
    return np.ascontiguousarray(np.linalg.lstsq(A, Y, rcond=None)[0])


def _window(k, dlogk):
    w = round(float(dlogk / (np.log(k[1]) - np.log(k[0]))))
    return int(w + (w + 1) % 2)


def _run(*, y, k, param, bias=1.0, window=0, mu=None, ell=0, want_xi=False, power_col=None, lib=None, device=0):
    lib = lib or _cabi.default_library()
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    if y.ndim != 2 or y.shape[1] != 20:
        raise ValueError("y must be [num_k, 20] (one output time of evolve_perturbations)")
    coef = _sg_weights(window) if window > 0 else None
    return lib.spectra_host(y, np.ascontiguousarray(k, dtype=np.float64), float(param["A_s"]), float(param["n_s"]), float(param["k_p"]),
                            float(bias), coef, mu, int(ell), want_xi, device=device)


def get_power_smoothed(*, k, y, dlogk: float, idx: int, param, lib=None):
    """Savitzky-Golay smoothed power spectrum in log-log space; the half-windows at both ends keep the raw signal."""
    if idx not in (4, 5):
        # the library smooths delta_m (4) and theta_m (5), what power_Kaiser needs; other columns go through column 4
        y = np.array(y, dtype=np.float64, copy=True)
        y[:, 4] = y[:, idx]
        idx = 4
    out = _run(y=y, k=np.asarray(k), param=param, window=_window(np.asarray(k), dlogk), lib=lib)
    return out["Ps_delta"] if idx == 4 else out["Ps_theta"]


def power_Kaiser(*, y, kmodes, bias: float, mu_sampling: bool = True, smooth_dlogk: float = None, nmu: int, param, lib=None):
    """Anisotropic Kaiser spectrum ``(b delta_m - mu^2 theta_m)^2`` on ``nmu`` bins of mu (or of the angle)."""
    kmodes = np.asarray(kmodes)
    mu = np.linspace(-1, 1, nmu) if mu_sampling else np.cos(np.linspace(0, np.pi, nmu))
    window = 0 if smooth_dlogk is None else _window(kmodes, smooth_dlogk)
    out = _run(y=y, k=kmodes, param=param, bias=bias, window=window, mu=np.ascontiguousarray(mu), lib=lib)
    return out["Pkmu"], mu


def power_multipoles(*, y, kmodes, b: float, param, lib=None):
    """Monopole, quadrupole and hexadecapole of the Kaiser spectrum."""
    out = _run(y=y, k=np.asarray(kmodes), param=param, bias=b, lib=lib)
    return out["P0"], out["P2"], out["P4"]


def get_xi_from_P(*, k, Pk, N: int = None, ell: int = 0, lib=None):
    """Correlation-function multipole from P(k) on a log-spaced grid by FFTlog (Talman 1978, Hamilton 2000).
    Returns ``(xi, r)`` with r = 2 pi / k in ascending order."""
    k = np.asarray(k, dtype=np.float64)
    Pk = np.asarray(Pk, dtype=np.float64)
    # the library transforms P_delta = amp(k)^2 y_4^2: feed y_4 = sqrt(Pk) with unit amplitude (A_s = 1/(2 pi^2), n_s = 1, k^-3 undone)
    y = np.zeros((k.size, 20))
    y[:, 4] = np.sqrt(Pk * k ** 3)
    param = dict(A_s=1.0 / (2 * np.pi ** 2), n_s=1.0, k_p=1.0)
    out = _run(y=y, k=k, param=param, ell=ell, want_xi=True, lib=lib)
    return out["xi"], out["r"]
