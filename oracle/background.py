"""CPU table producer for the hot path's INPUTS (test infrastructure, not product).

The Einstein-Boltzmann hot path consumes a ``param`` dict of scalars and natural
cubic splines that the reference builds upstream of the path
(``/root/reference/src/discoeb/background.py:140-342`` and
``thermodynamics_recfast.py:124-501``).  Neither JAX nor diffrax exist in this
environment, so this module restates that producer with NumPy/SciPy so that the
oracle and the CUDA path can be fed the same, physically realistic tables.

Parity scope: the tables are *inputs* shared by the oracle and the CUDA path, so
nothing here sits on the parity chain.  The ionisation history is integrated
with ``scipy.integrate.solve_ivp`` (the reference uses its own GRKT4 stepper,
``thermodynamics_recfast.py:283-300``); it is pinned against the reference's own
golden ``tests/resources/RECFAST_DISCO_EB_data.json`` (copied to
``tests/golden/``) at the reference's tolerance of 0.5 % (``tests/test_background.py:60-75``).

Only ``tests/``, ``tools/`` fixture generators, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline leg may import this file.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

__all__ = ["Spline", "nu_momentum_bins", "nu_background", "aprimeoa",
           "setup_background", "evolve_background", "fiducial_param"]


# --------------------------------------------------------------------------------------
# natural cubic spline  (spline_interpolation.py:8-113 ctor, :130-153 evaluate)
# --------------------------------------------------------------------------------------
class Spline:
    """Natural cubic spline in the (x, y, S=second derivatives) form of the reference."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64).copy()
        y = np.asarray(y, dtype=np.float64).copy()
        n = x.shape[0]
        h = np.diff(x)
        S = np.zeros(n)
        m = n - 2
        if m > 0:
            # Thomas algorithm on the interior second derivatives (ctor :49-85)
            lo = h[:-1]
            di = 2.0 * (h[:-1] + h[1:])
            up = h[1:]
            d = 6.0 * ((y[2:] - y[1:-1]) / h[1:] - (y[1:-1] - y[:-2]) / h[:-1])
            cp = np.zeros(m)
            dp = np.zeros(m)
            cp[0] = up[0] / di[0]
            dp[0] = d[0] / di[0]
            for i in range(1, m):
                den = di[i] - lo[i] * cp[i - 1]
                cp[i] = up[i] / den if i < m - 1 else 0.0
                dp[i] = (d[i] - lo[i] * dp[i - 1]) / den
            Si = np.zeros(m)
            Si[m - 1] = dp[m - 1]
            for i in range(m - 2, -1, -1):
                Si[i] = dp[i] - cp[i] * Si[i + 1]
            S[1:-1] = Si
        self.x, self.y, self.S = x, y, S

    def evaluate(self, xn):
        """spline_interpolation.py:130-153 (cubic extrapolation outside the knots)."""
        x, y, S = self.x, self.y, self.S
        n = x.shape[0]
        xr = np.real(xn)      # (complex xn / complex tables: the tangent oracle's complex step; index from real parts)
        idx = np.clip(np.searchsorted(np.real(x), xr, side="left") - 1, 0, n - 2)
        hl = x[idx + 1] - x[idx]
        t = (xn - x[idx]) / hl
        A = 1 - t
        B = t
        return (A * y[idx] + B * y[idx + 1]
                + ((A ** 3 - A) * S[idx] + (B ** 3 - B) * S[idx + 1]) * (hl ** 2) / 6.0)

    def derivative(self, xn):
        """spline_interpolation.py:191-212."""
        x, y, S = self.x, self.y, self.S
        n = x.shape[0]
        idx = np.clip(np.searchsorted(x, xn, side="left") - 1, 0, n - 2)
        hl = x[idx + 1] - x[idx]
        d = xn - x[idx]
        b = (y[idx + 1] - y[idx]) / hl - hl * (S[idx + 1] + 2 * S[idx]) / 6.0
        c = S[idx] / 2.0
        dd = (S[idx + 1] - S[idx]) / (6.0 * hl)
        return b + 2 * c * d + 3 * dd * d ** 2


# --------------------------------------------------------------------------------------
# massive-neutrino momentum quadrature (background.py:13-44, util.py:82-123)
# --------------------------------------------------------------------------------------
_FD_CONST = 5.682196976983475  # 7 pi^4 / 120 (background.py:23)

_CAMB_RULES = {
    3: ([0.913201, 3.37517, 7.79184], [0.0687359, 3.31435, 2.29911]),
    4: ([0.7, 2.62814, 5.90428, 12.0], [0.0200251, 1.84539, 3.52736, 0.289427]),
    5: ([0.583165, 2.0, 4.0, 7.26582, 13.0], [0.0081201, 0.689407, 2.8063, 2.05156, 0.12681]),
}


def nu_momentum_bins(nq: int):
    """q-nodes and kernel weights; CAMB 3/4/5-point rules else generalised Gauss-Laguerre."""
    if nq in _CAMB_RULES:
        q = np.array(_CAMB_RULES[nq][0])
        dlf = -q / (1.0 + np.exp(-q))
        w = np.array(_CAMB_RULES[nq][1]) / (-0.25 * dlf)
    else:
        alpha = 1
        i = np.arange(1, nq + 1)
        diag = 2.0 * i - 1.0 + alpha
        io = np.arange(1, nq)
        off = np.sqrt(io * (io + alpha))
        Jm = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
        q, V = np.linalg.eigh(Jm)
        w = V[0, :] ** 2 * 1.0  # Gamma(alpha+1) = 1! = 1
        w = w * q ** 3 / (1.0 + np.exp(-q)) * q ** (-alpha)
    return q, w / _FD_CONST


def nu_background(a, amnu, nq: int = 8):
    """rho, p, pseudo-p of one massive flavour in units of one massless flavour (background.py:47-68)."""
    q, w = nu_momentum_bins(nq)
    a = np.asarray(a, dtype=np.float64)[..., None]
    v = 1.0 / np.sqrt(1.0 + (a * amnu / q) ** 2)
    return (w / v).sum(-1), (w * v / 3).sum(-1), (w * v ** 3 / 3).sum(-1)


def _rho_de(a, p):
    return a ** (-3 * (1 + p["w_DE_0"] + p["w_DE_a"])) * np.exp(3 * (a - 1) * p["w_DE_a"])


def aprimeoa(p, a):
    """Conformal Hubble rate (background.py:100-122)."""
    rhonu = np.exp(p["logrhonu_of_loga_spline"].evaluate(np.log(a)))
    grho = (p["grhom"] * p["Omegam"] / a
            + (p["grhog"] + p["grhor"] * (p["Neff"] + p["Nmnu"] * rhonu)) / a ** 2
            + p["grhom"] * p["OmegaDE"] * _rho_de(a, p) * a ** 2
            + p["grhom"] * p["Omegak"])
    return np.sqrt(grho / 3.0)


def _dtauda(a, p):
    """background.py:71-80."""
    rhonu = np.exp(p["logrhonu_of_loga_spline"].evaluate(np.log(a)))
    g2 = (p["grhom"] * p["Omegam"] * a + (p["grhog"] + p["grhor"] * (p["Neff"] + p["Nmnu"] * rhonu))
          + p["grhom"] * p["OmegaDE"] * _rho_de(a, p) * a ** 4 + p["grhom"] * p["Omegak"] * a ** 2)
    return np.sqrt(3.0 / g2)


def _romb(f, lo, hi, divmax=6):
    """Romberg on 2**divmax+1 samples (stands in for jax_cosmo.scipy.integrate.romb)."""
    from scipy.integrate import romb as _sromb
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    ns = 2 ** divmax + 1
    u = np.linspace(0.0, 1.0, ns)
    xs = lo[..., None] + (hi - lo)[..., None] * u
    return _sromb(f(xs), dx=1.0 / (ns - 1), axis=-1) * (hi - lo)


def setup_background(p, amin=1e-9, amax=1.01):
    """Densities, neutrino splines, OmegaDE closure, taumin (background.py:140-188)."""
    c2ok = 1.62581581e4
    p["amin"], p["amax"] = amin, amax
    p["grhom"] = 3.33795017e-11 * p["H0"] ** 2
    p["grhog"] = 1.49594245e-13 * p["Tcmb"] ** 4
    p["grhor"] = 3.39739477e-14 * p["Tcmb"] ** 4
    p["adotrad"] = np.sqrt((p["grhog"] + p["grhor"] * (p["Neff"] + p["Nmnu"])) / 3.0)
    p["amnu"] = p["mnu"] * c2ok / p["Tcmb"]
    a = np.geomspace(amin * 0.9, amax * 1.1, 512)
    la = np.log(a)
    rho, pr, pp = nu_background(a, p["amnu"])
    p["logrhonu_of_loga_spline"] = Spline(la, np.log(rho))
    p["logpnu_of_loga_spline"] = Spline(la, np.log(pr))
    rhonu0 = np.exp(p["logrhonu_of_loga_spline"].evaluate(0.0))
    p["Omegamnu"] = p["grhor"] * rhonu0 / p["grhom"]
    Omegar = (p["Neff"] + p["Nmnu"] * rhonu0) * p["grhor"] / p["grhom"]
    Omegag = p["grhog"] / p["grhom"]
    p["OmegaDE"] = 1.0 - p["Omegak"] - Omegar - Omegag - p["Omegam"]
    p["taumin"] = amin / p["adotrad"]
    p["taumax"] = p["taumin"] + float(_romb(lambda x: _dtauda(x, p), amin, amax))
    return p


# --------------------------------------------------------------------------------------
# RECFAST ionisation history (physics of Seager, Sasselov & Scott 1999/2000, Wong+ 2008;
# arrangement follows thermodynamics_recfast.py:124-501)
# --------------------------------------------------------------------------------------
_G, _mH, _me, _mHe = 6.67430e-11, 1.67353284e-27, 9.1093837015e-31, 3.97146570884
_c, _hP, _kB, _sigT = 2.99792458e8, 6.62607015e-34, 1.380649e-23, 6.6524587321e-29
_arad = 4 * 5.670374419e-8 / _c
_bigH = 3.2407792902755102e-18
_densfac = 11.223810928601939
_E_He2 = 6.314878282674e5
_Lam_H, _Lam_He = 8.2245809, 51.3
_L_Hion, _L_Ha, _L_He1, _L_He2 = 1.096787737e7, 8.225916453e6, 1.98310772e7, 4.389088863e7
_L_He2s, _L_He2p = 1.66277434e7, 1.71134891e7
_A2Ps, _A2Pt = 1.798287e9, 177.58
_L_He2Pt, _L_He2St, _L_He2St_ion = 1.690871466e7, 1.5985597526e7, 3.8454693845e6
_sig2Ps, _sig2Pt = 1.436289e-22, 1.484872e-22
_AG1, _AG2, _zG1, _zG2, _wG1, _wG2 = -0.14, 0.05, 7.28, 6.75, 0.18, 0.33
_Hfrac = 1e-3
_lam_a, _lam_aHe = 1.0 / _L_Ha, 1.0 / _L_He2p
_CDB = _hP * _c * (_L_Hion - _L_Ha) / _kB
_CDB_He = _hP * _c * (_L_He1 - _L_He2s) / _kB
_CB1 = _hP * _c * _L_Hion / _kB
_CB1_He1 = _hP * _c * _L_He1 / _kB
_CR = 2.0 * np.pi * (_me / _hP) * (_kB / _hP)
_CK = _lam_a ** 3 / (8.0 * np.pi)
_CK_He = _lam_aHe ** 3 / (8.0 * np.pi)
_CL = _c * _hP / (_kB * _lam_a)
_CL_He = _c * _hP / (_kB / _L_He2s)
_CT = (8.0 / 3.0) * (_sigT / (_me * _c)) * _arad
_Bfact = _hP * _c * (_L_He2p - _L_He2s) / _kB
_CL_PSt = _hP * _c * (_L_He2Pt - _L_He2St) / _kB
_hcL2St = _hP * _c * _L_He2St / _kB
_aPPB, _bPPB, _cPPB, _dPPB = 4.309, -0.6166, 0.6703, 0.5300
_aVF, _bVF, _T0, _T1 = 10.0 ** (-16.744), 0.711, 10 ** 0.477121, 10 ** 5.114
_atrip, _btrip = 10.0 ** (-16.306), 0.761


def _ion_rhs(a, y, p):
    """d(x_H, x_He, T_m)/da  (thermodynamics_recfast.py:124-281)."""
    z1 = 1.0 / a
    hh = p["H0"] / 100.0
    HO = hh * _bigH
    fu = 1.105
    fHe = p["fHe"]
    Nnow = _densfac * hh * hh * p["Omegab"] * (1.0 - p["YHe"])
    xH, xHe, Tm = y[0], y[1], abs(y[2])
    x = xH + fHe * xHe
    n = Nnow * z1 ** 3
    nHe = fHe * n
    Tr = p["Tcmb"] * z1
    Hz = (1e-5 * float(aprimeoa(p, a))) / a * _c * _bigH

    T4 = Tm / 1e4
    crt = (_CR * Tm) ** 1.5
    Rdn = 1e-19 * _aPPB * T4 ** _bPPB / (1 + _cPPB * T4 ** _dPPB)
    Rup = Rdn * crt * np.exp(-_CDB / Tm)
    s0, s1 = np.sqrt(Tm / _T0), np.sqrt(Tm / _T1)
    RdnHe = _aVF / (s0 * (1 + s0) ** (1 - _bVF) * (1 + s1) ** (1 + _bVF))
    RupHe = 4 * RdnHe * crt * np.exp(-_CDB_He / Tm)
    HeB = np.exp(min(680.0, _Bfact / Tm))
    Rdn_t = _atrip / (s0 * (1 + s0) ** (1 - _btrip) * (1 + s1) ** (1 + _btrip))
    Rup_t = Rdn_t * np.exp(-_hP * _c * _L_He2St_ion / (_kB * Tm)) * crt * (4 / 3)

    lz = np.log(z1)
    K = _CK / Hz * (1 + _AG1 * np.exp(-((lz - _zG1) / _wG1) ** 2) + _AG2 * np.exp(-((lz - _zG2) / _wG2) ** 2))

    omHe, omH = 1 - xHe, 1 - xH
    nHe1 = nHe * omHe
    with np.errstate(all="ignore"):
        tauHe_s = _A2Ps * _CK_He * 3 * nHe1 / Hz
        pHe_s = (1 - np.exp(-tauHe_s)) / tauHe_s
        dop = np.sqrt(2 * _kB * Tm / (_mH * _mHe * _c ** 2))
        g2Ps = (3 * _A2Ps * fHe * omHe * _c ** 2) / (np.sqrt(np.pi) * _sig2Ps * 8 * np.pi * (_c * _L_He2p * dop) * omH * (_c * _L_He2p) ** 2)
        AHcon = _A2Ps / (1 + 0.36 * g2Ps ** 0.86)
        if xHe < 5e-9 or xHe > 0.98:
            KHe = _CK_He / Hz
        elif xH < 0.9999999:
            KHe = 1.0 / ((_A2Ps * pHe_s + AHcon) * 3.0 * nHe1)
        else:
            KHe = 1.0 / (_A2Ps * pHe_s * 3.0 * nHe1)
        tauHe_t = _A2Pt * nHe1 * 3 / (8 * np.pi * Hz * _L_He2Pt ** 3)
        pHe_t = (1 - np.exp(-tauHe_t)) / tauHe_t
        g2Pt = (3 * _A2Pt * fHe * omHe * _c ** 2) / (np.sqrt(np.pi) * _sig2Pt * 8 * np.pi * (_c * _L_He2Pt * dop) * omH * (_c * _L_He2Pt) ** 2)
        AHcon_t = _A2Pt / (1 + 0.66 * g2Pt ** 0.9) / 3
        eps = np.exp(-_CL_PSt / Tm)
        Cf = _A2Pt * pHe_t * eps if xH > 0.99999 else (_A2Pt * pHe_t + AHcon_t) * eps
        Cf = Cf / (Rup_t + Cf)

    timeTh = (1 / (_CT * Tr ** 4)) * (1 + x + fHe) / x
    timeH = 2 / (3 * HO * z1 ** 1.5)
    Hzz = Hz * z1

    rd = x * xH * n * Rdn - Rup * omH * np.exp(-_CL / Tm)
    if xH > 0.99:
        f0 = 0.0
    elif xH > 0.985:
        f0 = rd / Hzz
    else:
        KL = K * _Lam_H * n * omH
        f0 = rd * (1 + KL) / (Hzz * (1 / fu + KL / fu + K * Rup * n * omH))

    if xHe < 1e-8:
        f1 = 0.0
    else:
        rdHe = x * xHe * n * RdnHe - RupHe * omHe * np.exp(-_CL_He / Tm)
        KLHe = KHe * _Lam_He * nHe1 * HeB
        f1 = rdHe * (1 + KLHe) / (Hzz * (1 + KHe * (_Lam_He + RupHe) * nHe1 * HeB))
        if not (xHe < 5e-9 or xHe > 0.98):
            tr = x * xHe * n * Rdn_t - omHe * 3 * Rup_t * np.exp(-_hcL2St / Tm)
            f1 += tr * Cf / Hzz

    if timeTh < _Hfrac * timeH:
        f2 = Tm / z1
    else:
        f2 = _CT * Tr ** 4 * x / (1 + x + fHe) * (Tm - Tr) / Hzz + 2 * Tm / z1
    return np.array([f0, f1, f2]) * (-1.0 / a ** 2)


def _saha_HeII(a, p):
    """thermodynamics_recfast.py:302-321."""
    T = p["Tcmb"] / a
    fHe = p["fHe"]
    Hfac = 1 / (1.0e6 * 3.0856775807e13)
    nH = 3 * (p["H0"] * Hfac) ** 2 / (8 * np.pi * _G) * p["Omegab"] / (_mH / (1 - p["YHe"])) / a ** 3
    A, B = 1 + fHe, 1 + 2 * fHe
    with np.errstate(over="ignore"):
        R = (2 * np.pi * _me * _kB / _hP ** 2 * T) ** 1.5 / nH * np.exp(-_E_He2 / T)
    big = R > 1e5
    Rs = np.where(big, R, 1.0)
    hi = fHe * (1 - B / Rs + (1 + 5 * fHe + 6 * fHe ** 2) / Rs ** 2)
    Rl = np.where(big, 0.0, R)
    lo = -(Rl - A) / 2 + np.sqrt((Rl - A) ** 2 / 4 + Rl * B) - A
    return np.where(big, hi, lo)


def _adaptive_a_grid(a0, a1, N):
    """thermodynamics_recfast.py:324-360."""
    n1 = max(8, int(N * 0.05))
    n2 = max(8, int(N * 0.10))
    n3 = int(N * 0.50)
    n4 = N - n1 - n2 - n3
    b1 = max(1.0 / 3001.0, a0)
    b2, b3 = 1.0 / 1401.0, 1.0 / 601.0
    return np.concatenate([np.geomspace(a0, b1, n1, endpoint=False), np.geomspace(b1, b2, n2, endpoint=False),
                           np.geomspace(b2, b3, n3, endpoint=False), np.geomspace(b3, a1, n4)])


def _thermal_history(p, N, rtol=1e-8, atol=1e-11):
    """Interval-by-interval integration with Saha switches (thermodynamics_recfast.py:362-452)."""
    a = _adaptive_a_grid(p["amin"], p["amax"], N + 1)
    out = np.zeros((6, N))
    hh = p["H0"] / 100.0
    Nnow = 3.0 * (hh * _bigH) ** 2 * p["Omegab"] / (8.0 * np.pi * _G * _mH / (1.0 - p["YHe"]))
    fHe = p["fHe"]
    Tc = p["Tcmb"]
    h = 1e-6

    def saha_He1(z1):
        rhs = np.exp(1.5 * np.log(_CR * Tc / z1) - _CB1_He1 / (Tc * z1)) / Nnow * 4.0
        return 0.5 * (np.sqrt((rhs - 1.0) ** 2 + 4.0 * (1.0 + fHe) * rhs) - (rhs - 1.0))

    def saha_H(z1):
        rhs = np.exp(1.5 * np.log(_CR * Tc / z1) - _CB1 / (Tc * z1)) / Nnow
        return 0.5 * (np.sqrt(rhs ** 2 + 4.0 * rhs) - rhs)

    def integrate(a0, a1, y0):
        sol = solve_ivp(lambda t, y: _ion_rhs(t, y, p), (a0, a1), y0, method="LSODA", rtol=rtol, atol=atol)
        ye = sol.y[:, -1]
        return np.concatenate([ye, _ion_rhs(a1, ye, p)])

    for i in range(N):
        a0, a1 = a[i], a[i + 1]
        z1e = 1.0 / a1
        prev = out[:, i - 1] if i > 0 else None
        if z1e - 1.0 > 3500.0:
            out[:, i] = [1.0, 1.0, Tc * z1e, 0.0, 0.0, -Tc * z1e]
        elif i > 0 and prev[1] > 0.99:
            # helium singly ionised by Saha; d/dz by a centred difference of the closed form
            # (the reference differentiates it symbolically, :411-416)
            x0 = saha_He1(z1e)
            dxdz = (saha_He1(z1e * (1 + h)) - saha_He1(z1e * (1 - h))) / (2 * h * z1e) / fHe
            out[:, i] = [1.0, (x0 - 1.0) / fHe, Tc * z1e, 0.0, dxdz * (-1.0 / a1 ** 2), -Tc * z1e]
        elif i > 0 and prev[0] > 0.99:
            x0 = saha_H(z1e)
            dxdz = (saha_H(z1e * (1 + h)) - saha_H(z1e * (1 - h))) / (2 * h * z1e)
            ys = integrate(a0, a1, prev[:3])
            ys[0] = x0
            ys[3] = dxdz * (-1.0 / a1 ** 2)
            out[:, i] = ys
        else:
            y0 = prev[:3] if i > 0 else np.array([1.0, 1.0, Tc / a0])
            out[:, i] = integrate(a0, a1, y0)
    return out, a[1:]


def evolve_background(p, num_thermo: int = 256):
    """Produce every table the hot path reads (background.py:191-255 with thermo_module='RECFAST')."""
    p = setup_background(dict(p))
    p["fHe"] = p["YHe"] / (_mHe * (1.0 - p["YHe"]))
    y, a = _thermal_history(p, num_thermo)
    xHeII = _saha_HeII(a, p)
    xe = y[0] + p["fHe"] * y[1] + xHeII
    mu = 1 / (1 + (1 / _mHe - 1) * p["YHe"] + (1 - p["YHe"]) * xe)
    Tm = y[2]
    daTmda = Tm + a * y[5]
    cs2 = _kB / _mH / _c ** 2 / mu * Tm * (4 - daTmda / Tm) / 3
    dtau = _romb(lambda x: _dtauda(x, p), a[:-1], a[1:])
    tau = np.concatenate([[p["taumin"]], p["taumin"] + np.cumsum(dtau)])
    p["aexp"], p["tau"], p["xe"], p["cs2"], p["Tm"] = a, tau, xe, cs2, Tm
    p["tau_of_a_spline"] = Spline(a, tau)
    p["a_of_tau_spline"] = Spline(tau, a)
    p["xe_of_tau_spline"] = Spline(tau, xe)
    p["xe_of_loga_spline"] = Spline(np.log(a), xe)
    p["cs2a_of_loga_spline"] = Spline(np.log(a), a * cs2)
    return p


def fiducial_param(**over):
    """Fiducial cosmology of the reference's tests (tests/test_perturbations.py:11-57)."""
    Tnu = (4 / 11) ** (1 / 3)
    Nmass = 1
    p = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0,
             A_s=2.1064e-09, n_s=0.96822, k_p=0.05, H0=100 * 0.67742, Tcmb=2.7255, YHe=0.248,
             Neff=3.046 - Nmass * (Tnu / ((4 / 11) ** (1 / 3))) ** 4, Nmnu=Nmass, mnu=0.06)
    p.update(over)
    return p
