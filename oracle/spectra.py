"""CPU ORACLE (NumPy) of the spectra post-processing entry points of the reference's host API
(``/root/reference/src/discoeb/perturbations.py:1063-1224``) -- TEST INFRASTRUCTURE: a restatement of those pure array
functions, pinned against SciPy (tests/test_spectra_cpu.py); the product computes them on the GPU
(disco-eb_b200/csrc/deb_spectra.cu behind disco-eb_b200/discoeb_b200/spectra.py) and is compared with this module.
Only tests/ may import it.

=========================  ============================================
``get_power``               perturbations.py:1101-1123
``get_power_smoothed``      perturbations.py:1126-1160  (util.savgol_filter :407-444)
``power_Kaiser``            perturbations.py:1162-1199
``power_multipoles``        perturbations.py:1202-1224
``get_xi_from_P``           perturbations.py:1063-1098  (FFTlog; util.lngamma_complex_e :12-44)
=========================  ============================================
"""
from __future__ import annotations

import numpy as np

__all__ = ["get_power", "get_power_smoothed", "power_Kaiser", "power_multipoles", "get_xi_from_P", "savgol_filter",
           "lngamma_complex_e"]


def get_power(*, k, y, idx: int, param):
    """``2 pi^2 A_s (k/k_p)^(n_s-1) k^-3 y[..., idx]^2``."""
    k = np.asarray(k)
    y = np.asarray(y)
    return 2 * np.pi ** 2 * param["A_s"] * (k / param["k_p"]) ** (param["n_s"] - 1) * k ** (-3) * y[..., idx] ** 2


def savgol_filter(*, y, window_length: int, polyorder: int):
    """Savitzky-Golay smoothing as the reference defines it: least-squares centre weights of a degree-``polyorder``
    fit over ``window_length`` samples, applied by a zero-padded 'same' convolution (util.py:407-444)."""
    window_length = int(window_length)
    halflen, rem = divmod(window_length, 2)
    pos = halflen - 0.5 if rem == 0 else halflen
    x = np.arange(-pos, window_length - pos, dtype=float)[::-1]
    order = np.arange(polyorder + 1).reshape(-1, 1)
    A = x ** order
    Y = np.zeros(polyorder + 1)
    Y[0] = 1.0
    coeffs = np.linalg.lstsq(A, Y, rcond=None)[0]
    return np.convolve(np.asarray(y), coeffs, mode="same")


def get_power_smoothed(*, k, y, dlogk: float, idx: int, param):
    """Savitzky-Golay smoothed power spectrum in log-log space; the half-windows at both ends keep the raw signal."""
    k = np.asarray(k)
    window_length = round(float(dlogk / (np.log(k[1]) - np.log(k[0]))))
    window_length += (window_length + 1) % 2
    Pm = get_power(y=y, k=k, idx=idx, param=param)
    Pms = np.exp(savgol_filter(y=np.log(Pm), window_length=window_length, polyorder=3))
    h = window_length // 2
    if h > 0:
        Pms[:h] = Pm[:h]
    # (the reference writes Pms.at[-window_length//2:], i.e. floor(-w/2) = -(h+1) for the odd w it constructs)
    Pms[-window_length // 2:] = Pm[-window_length // 2:]
    return Pms


def power_Kaiser(*, y, kmodes, bias: float, mu_sampling: bool = True, smooth_dlogk: float = None, nmu: int, param):
    """Anisotropic Kaiser spectrum ``(b delta_m - mu^2 theta_m)^2`` on ``nmu`` bins of mu (or of the angle)."""
    y = np.asarray(y)
    kmodes = np.asarray(kmodes)
    mu = np.linspace(-1, 1, nmu) if mu_sampling else np.cos(np.linspace(0, np.pi, nmu))
    if smooth_dlogk is None:
        amp = np.sqrt(2 * np.pi ** 2 * param["A_s"] * (kmodes / param["k_p"]) ** (param["n_s"] - 1) * kmodes ** (-3))
        deltam, thetam = amp * y[:, 4], amp * y[:, 5]
    else:
        deltam = np.sqrt(get_power_smoothed(y=y, k=kmodes, dlogk=smooth_dlogk, idx=4, param=param))
        thetam = -np.sqrt(get_power_smoothed(y=y, k=kmodes, dlogk=smooth_dlogk, idx=5, param=param))
    return (bias * deltam[:, None] - mu[None, :] ** 2 * thetam[:, None]) ** 2, mu


def power_multipoles(*, y, kmodes, b: float, param):
    """Monopole, quadrupole and hexadecapole of the Kaiser spectrum."""
    y = np.asarray(y)
    kmodes = np.asarray(kmodes)
    amp = np.sqrt(2 * np.pi ** 2 * param["A_s"] * (kmodes / param["k_p"]) ** (param["n_s"] - 1) * kmodes ** (-3))
    deltam, thetam = amp * y[:, 4], amp * y[:, 5]
    P0 = b ** 2 * deltam ** 2 - 2 * b / 3 * deltam * thetam + 1 / 5 * thetam ** 2
    P2 = -4 * b / 3 * deltam * thetam + 4 / 7 * thetam ** 2
    P4 = 8 / 35 * thetam ** 2
    return P0, P2, P4


_LANCZOS_7 = np.array([0.99999999999980993227684700473478, 676.520368121885098567009190444019,
                       -1259.13921672240287047156078755283, 771.3234287776530788486528258894,
                       -176.61502916214059906584551354, 12.507343278686904814458936853,
                       -0.13857109526572011689554707, 9.984369578019570859563e-6, 1.50563273514931155834e-7])


def lngamma_complex_e(z):
    """log Gamma(z) for complex z by the Lanczos method with reflection for Re z <= 1/2 (util.py:12-44, after GSL);
    vectorised over z."""
    z = np.asarray(z, dtype=np.complex128)

    def lanczos(zz):
        zz = zz - 1.0
        t = zz[..., None] + np.arange(1, 9)
        Ag = _LANCZOS_7[0] + np.sum(_LANCZOS_7[1:] / np.abs(t) ** 2 * np.conj(t), -1)
        return (zz + 0.5) * np.log(zz + 7.5) - (zz + 7.5) + 0.9189385332046727418 + np.log(Ag)

    refl = np.real(z) <= 0.5
    out = np.empty_like(z)
    if np.any(~refl):
        out[~refl] = lanczos(z[~refl])
    if np.any(refl):
        zr = z[refl]
        out[refl] = 1.14472988584940017414342735135 - np.log(np.sin(np.pi * zr)) - lanczos(1.0 - zr)
    return out


def get_xi_from_P(*, k, Pk, N: int = None, ell: int = 0):
    """Correlation-function multipole from P(k) on a log-spaced grid by FFTlog (Talman 1978, Hamilton 2000).
    Returns ``(xi, r)`` with r = 2 pi / k in ascending order."""
    k = np.asarray(k)
    Pk = np.asarray(Pk)
    N = len(k)
    L = np.log(k[N - 1] / k[0])
    fPk = np.fft.rfft(Pk * k ** 1.5)
    ki = np.pi * np.arange(N // 2 + 1) / L
    zp = (1.5 + ell) / 2 + 1j * ki
    theta = np.imag(lngamma_complex_e(zp))
    fPk = fPk * np.exp(2j * (theta - np.log(np.pi) * ki))
    r = 2 * np.pi / k
    xi = np.real(1j ** ell * np.fft.irfft(fPk, n=N) / (2 * np.pi * r) ** 1.5)
    return xi[::-1], r[::-1]
