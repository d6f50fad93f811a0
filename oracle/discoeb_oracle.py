"""CPU ORACLE (NumPy/SciPy, float64) for the Einstein-Boltzmann hot path -- TEST INFRASTRUCTURE.

Restates, function by function, the reference's algorithm for the path
``evolve_perturbations -> evolve_one_mode -> diffeqsolve(Rodas5Transformed) -> get_power``
(citations are into ``/root/reference/src/discoeb/``):

=========================  ==================================================
here                        reference
=========================  ==================================================
``rhs``                     perturbations.py:84-371 (+ nu_perturb :22-50)
``jacobian``, ``dfdt``      what ``jax.jacfwd`` yields at ode_integrators_stiff.py:772, :779
``rodas5_step``             ode_integrators_stiff.py:620-688, :718-834
``scaled_error_norm``       perturbations.py:701-711, :759 + diffrax PIDController
``integrate_modes``         diffrax.diffeqsolve loop as called at perturbations.py:751-767, :770
``determine_starting_time`` perturbations.py:630-681, util.py:365-396
``adiabatic_ics``           perturbations.py:526-627
``convert_to_output``       perturbations.py:374-523
``evolve_perturbations``    perturbations.py:728-781, :926-997
``get_power``               perturbations.py:1101-1123
=========================  ==================================================

Third-party arithmetic that is NOT under /root/reference: ``diffrax`` (unpinned in the
reference's pyproject; a comment targets 0.7.0) supplies the adaptive loop, ``PIDController``,
``SaveAt(ts)`` and ``LocalLinearInterpolation``; ``jax.scipy.linalg.lu_factor/lu_solve`` the dense
pivoted LU.  Their published semantics are restated in ``integrate_modes``
and the dense solves go through LAPACK ``getrf/getrs`` (``scipy.linalg``), like jaxlib on CPU.

PARITY UNPINNED at the 1e-5 level: neither JAX nor diffrax can be installed here, so this
restatement cannot be run against the real reference.  What pins it: (i) the reference's own
acceptance test -- P_bc(k) within 0.5 % of the stored CLASS curve at z=99
(tests/test_perturbations.py:128-142, tests/resources/CLASS_data.json) -- re-run against this
oracle in tests/test_oracle_golden.py; (ii) the analytic Jacobian here is checked against a
brute-force Jacobian of the restated RHS (linearity + complex step), i.e. what jacfwd computes.
``tools/crosscheck_jax.py`` runs the real reference beside it wherever JAX exists.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product path (disco-eb_b200/) never does.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .background import nu_momentum_bins, aprimeoa

# Rodas5 coefficients in transformed form (ode_integrators_stiff.py:622-687)
GAMMA = 0.19
A = {(2, 1): 2.0,
     (3, 1): 3.040894194418781, (3, 2): 1.041747909077569,
     (4, 1): 2.576417536461461, (4, 2): 1.622083060776640, (4, 3): -0.9089668560264532,
     (5, 1): 2.760842080225597, (5, 2): 1.446624659844071, (5, 3): -0.3036980084553738, (5, 4): 0.2877498600325443,
     (6, 1): -14.09640773051259, (6, 2): 6.925207756232704, (6, 3): -41.47510893210728, (6, 4): 2.343771018586405,
     (6, 5): 24.13215229196062}
C = {(2, 1): -10.31323885133993,
     (3, 1): -21.04823117650003, (3, 2): -7.234992135176716,
     (4, 1): 32.22751541853323, (4, 2): -4.943732386540191, (4, 3): 19.44922031041879,
     (5, 1): -20.69865579590063, (5, 2): -8.816374604402768, (5, 3): 1.260436877740897, (5, 4): -0.7495647613787146,
     (6, 1): -46.22004352711257, (6, 2): -17.49534862857472, (6, 3): -289.6389582892057, (6, 4): 93.60855400400906,
     (6, 5): 318.3822534212147,
     (7, 1): 34.20013733472935, (7, 2): -14.15535402717690, (7, 3): 57.82335640988400, (7, 4): 25.83362985412365,
     (7, 5): 1.408950972071624, (7, 6): -6.551835421242162,
     (8, 1): 42.57076742291101, (8, 2): -13.80770672017997, (8, 3): 93.98938432427124, (8, 4): 18.77919633714503,
     (8, 5): -31.58359187223370, (8, 6): -6.685968952921985, (8, 7): -5.810979938412932}
CT = {2: 0.38, 3: 0.3878509998321533, 4: 0.4839718937873840, 5: 0.4570477008819580}
D = {1: GAMMA, 2: -0.1823079225333714636, 3: -0.319231832186874912, 4: 0.3449828624725343, 5: -0.377417564392089818}

AKTHOM_RHS = 2.3038921003709498e-9     # perturbations.py:215
AKTHOM_START = 2.3048e-9               # perturbations.py:648 (the reference really uses a different constant)


class Dims:
    """State-vector layout (perturbations.py:115-119, :283-284, :316, :739)."""

    def __init__(self, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3):
        self.lmaxg, self.lmaxgp, self.lmaxr, self.lmaxnu, self.nq = lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax
        self.ig = 7
        self.igp = 7 + (lmaxg + 1)
        self.ir = 9 + lmaxg + lmaxgp
        self.iq0 = 10 + lmaxg + lmaxgp + lmaxr
        self.n = 7 + (lmaxg + 1) + (lmaxgp + 1) + (lmaxr + 1) + nqmax * (lmaxnu + 1) + 2
        q, w = nu_momentum_bins(nqmax)
        self.q, self.w = q, w
        self.dlfdlq = -q / (1.0 + np.exp(-q))


def _bg(p, a):
    """Background coefficients shared by the RHS and the Jacobian (perturbations.py:176-218)."""
    loga = np.log(a)
    b = {}
    b["cs2"] = p["cs2a_of_loga_spline"].evaluate(loga) / a
    b["xe"] = p["xe_of_loga_spline"].evaluate(loga)
    b["pb43"] = 4.0 / 3.0 * p["grhog"] / (p["grhom"] * p["Omegab"] * a)
    b["w_Q"] = p["w_DE_0"] + p["w_DE_a"] * (1.0 - a)
    b["rho_Q"] = a ** (-3 * (1 + p["w_DE_0"] + p["w_DE_a"])) * np.exp(3 * (a - 1) * p["w_DE_a"])
    b["H"] = aprimeoa(p, a)
    w_Q_prime = -p["w_DE_a"] * b["H"] * a
    b["ca2_Q"] = b["w_Q"] - w_Q_prime / 3 / ((1 + b["w_Q"]) + 1e-6) / b["H"]
    akthom = AKTHOM_RHS * (1.0 - p["YHe"]) * p["Omegab"] * p["H0"] ** 2
    b["opac"] = b["xe"] * akthom / a ** 2
    return b


def rhs(tau, y, p, k, d: Dims):
    """f(tau, y) for a batch of modes: y[..., n], tau[...], k[...] (perturbations.py:84-371)."""
    nq, iq0 = d.nq, d.iq0
    Omegac = p["Omegam"] - p["Omegab"]
    f = np.zeros_like(y)
    a = y[..., 0]
    eta = y[..., 2]
    deltac, thetac, deltab, thetab = y[..., 3], y[..., 4], y[..., 5], y[..., 6]
    deltag, thetag, shearg = y[..., 7], y[..., 8], y[..., 9] / 2.0
    deltar, thetar, shearr = y[..., d.ir], y[..., d.ir + 1], y[..., d.ir + 2] / 2.0
    deltaq, thetaq = y[..., -2], y[..., -1]
    b = _bg(p, a)
    cs2, pb43, w_Q, rho_Q, H, ca2_Q, opac = b["cs2"], b["pb43"], b["w_Q"], b["rho_Q"], b["H"], b["ca2_Q"], b["opac"]
    cs2_Q = p["cs2_DE"]
    rpt_Q = (1 + w_Q) * rho_Q * p["grhom"] * p["OmegaDE"] * thetaq * a ** 2
    k2 = k ** 2

    f[..., 0] = H * a
    # massive-nu moment sums (nu_perturb, :22-50)
    aq = a[..., None] * p["amnu"] / d.q
    v = 1 / np.sqrt(1 + aq ** 2)
    psi0, psi1, psi2 = y[..., iq0:iq0 + nq], y[..., iq0 + nq:iq0 + 2 * nq], y[..., iq0 + 2 * nq:iq0 + 3 * nq]
    drhonu = np.sum(d.w * psi0 / v, -1)
    dpnu = np.sum(d.w * psi0 * v, -1) / 3
    fnu = np.sum(d.w * psi1, -1)

    dgrho = (p["grhom"] * (Omegac * deltac + p["Omegab"] * deltab) / a
             + (p["grhog"] * deltag + p["grhor"] * (p["Neff"] * deltar + p["Nmnu"] * drhonu)) / a ** 2
             + p["grhom"] * p["OmegaDE"] * deltaq * rho_Q * a ** 2)
    dgpres = ((p["grhog"] * deltag + p["grhor"] * p["Neff"] * deltar) / a ** 2 / 3.0
              + p["grhor"] * p["Nmnu"] * dpnu / a ** 2
              + (cs2_Q * p["grhom"] * p["OmegaDE"] * deltaq * rho_Q * a ** 2
                 + (cs2_Q - ca2_Q) * (3 * H * rpt_Q / k2)))
    dgtheta = (p["grhom"] * (Omegac * thetac + p["Omegab"] * thetab) / a
               + 4.0 / 3.0 * (p["grhog"] * thetag + p["Neff"] * p["grhor"] * thetar) / a ** 2
               + p["Nmnu"] * p["grhor"] * k * fnu / a ** 2
               + rpt_Q)
    f[..., 1] = -(dgrho + 3.0 * dgpres) * a
    hprime = (2.0 * k2 * eta + dgrho) / H
    etaprime = 0.5 * dgtheta / k2
    alpha = (hprime + 6.0 * etaprime) / 2.0 / k2
    f[..., 2] = etaprime
    f[..., 3] = -thetac - 0.5 * hprime
    f[..., 4] = -H * thetac
    f[..., 5] = -thetab - 0.5 * hprime
    f[..., 6] = -H * thetab + k2 * cs2 * deltab + pb43 * opac * (thetag - thetab)

    ig, igp, ir = d.ig, d.igp, d.ir
    Lg, Lp, Lr, Ln = d.lmaxg, d.lmaxgp, d.lmaxr, d.lmaxnu
    polter = y[..., ig + 2] + y[..., igp] + y[..., igp + 2]
    f[..., ig] = 4.0 / 3.0 * (-thetag - 0.5 * hprime)
    f[..., ig + 1] = k2 * (0.25 * deltag - shearg) - opac * (thetag - thetab)
    f[..., ig + 2] = (8.0 / 15.0 * (thetag + k2 * alpha) - 3 / 5 * k * y[..., ig + 3]
                      - opac * (y[..., ig + 2] - 0.1 * polter))
    ell = np.arange(3, Lg)
    kk, op = k[..., None], opac[..., None]
    f[..., ig + ell] = kk / (2 * ell + 1) * (ell * y[..., ig + ell - 1] - (ell + 1) * y[..., ig + ell + 1]) - op * y[..., ig + ell]
    f[..., ig + Lg] = k * y[..., ig + Lg - 1] - (Lg + 1) / tau * y[..., ig + Lg] - opac * y[..., ig + Lg]
    # polarisation (:306-312); the l=0 row multiplies y[igp-1] by zero
    ell = np.arange(0, Lp)
    f[..., igp + ell] = kk / (2 * ell + 1) * (ell * y[..., igp + ell - 1] - (ell + 1) * y[..., igp + ell + 1]) - op * y[..., igp + ell]
    f[..., igp] += opac * polter / 2
    f[..., igp + 2] += opac * polter / 10
    f[..., igp + Lp] = k * y[..., igp + Lp - 1] - (Lp + 1) / tau * y[..., igp + Lp] - opac * y[..., igp + Lp]
    # massless neutrinos (:316-327)
    f[..., ir] = 4.0 / 3.0 * (-thetar - 0.5 * hprime)
    f[..., ir + 1] = k2 * (0.25 * deltar - shearr)
    f[..., ir + 2] = 8.0 / 15.0 * (thetar + k2 * alpha) - 0.6 * k * y[..., ir + 3]
    ell = np.arange(3, Lr)
    f[..., ir + ell] = kk / (2 * ell + 1) * (ell * y[..., ir + ell - 1] - (ell + 1) * y[..., ir + ell + 1])
    f[..., ir + Lr] = k * y[..., ir + Lr - 1] - (Lr + 1) / tau * y[..., ir + Lr]
    # massive neutrinos (:331-360), l-major layout iq0 + l*nq + i
    dl = d.dlfdlq
    psi3 = y[..., iq0 + 3 * nq:iq0 + 4 * nq]
    f[..., iq0:iq0 + nq] = -kk * v * psi1 + hprime[..., None] * dl / 6.0
    f[..., iq0 + nq:iq0 + 2 * nq] = kk * v * (psi0 - 2.0 * psi2) / 3.0
    f[..., iq0 + 2 * nq:iq0 + 3 * nq] = kk * v * (2 * psi1 - 3 * psi3) / 5.0 - (hprime / 15 + 2 / 5 * etaprime)[..., None] * dl
    for l in range(3, Ln):
        lo = y[..., iq0 + (l - 1) * nq:iq0 + l * nq]
        hi = y[..., iq0 + (l + 1) * nq:iq0 + (l + 2) * nq]
        f[..., iq0 + l * nq:iq0 + (l + 1) * nq] = kk * v / (2 * l + 1) * (l * lo - (l + 1) * hi)
    f[..., iq0 + Ln * nq:iq0 + (Ln + 1) * nq] = (kk * v * y[..., iq0 + (Ln - 1) * nq:iq0 + Ln * nq]
                                                - (Ln + 1) / tau[..., None] * y[..., iq0 + Ln * nq:iq0 + (Ln + 1) * nq])
    # dark-energy fluid (:364-369)
    f[..., -2] = (-(1 + w_Q) * (thetaq + 0.5 * hprime) - 3 * (cs2_Q - w_Q) * H * deltaq
                  - 9 * (1 + w_Q) * (cs2_Q - ca2_Q) * H ** 2 / k2 * thetaq)
    f[..., -1] = -(1 - 3 * cs2_Q) * H * thetaq + cs2_Q / (1 + w_Q) * k2 * deltaq
    return f


def _spl_d(sp, xn):
    """d/dxn of Spline.evaluate's own formula (what forward-mode AD of spline_interpolation.py:130-153 yields)."""
    x, yv, S = sp.x, sp.y, sp.S
    n = x.shape[0]
    idx = np.clip(np.searchsorted(np.real(x), np.real(xn), side="left") - 1, 0, n - 2)
    hl = x[idx + 1] - x[idx]
    t = (xn - x[idx]) / hl
    A, B = 1 - t, t
    return (yv[idx + 1] - yv[idx]) / hl + ((1 - 3 * A ** 2) * S[idx] + (3 * B ** 2 - 1) * S[idx + 1]) * hl / 6.0


def rhs_da(tau, y, p, k, d: Dims):
    """Hand-derived d f / d a at fixed y[1:] (the scale-factor column of the Jacobian).  Holomorphic in
    every input, so it can sit inside an outer complex step; checked against the complex-step column
    of ``jacobian`` on real inputs in tests/test_oracle_tangent.py."""
    nq, iq0 = d.nq, d.iq0
    ig, igp, ir = d.ig, d.igp, d.ir
    Lg, Lp, Lr, Ln = d.lmaxg, d.lmaxgp, d.lmaxr, d.lmaxnu
    Omegac = p["Omegam"] - p["Omegab"]
    a = y[..., 0]
    eta = y[..., 2]
    deltac, thetac, deltab, thetab = y[..., 3], y[..., 4], y[..., 5], y[..., 6]
    deltag, thetag = y[..., 7], y[..., 8]
    deltar, thetar = y[..., ir], y[..., ir + 1]
    deltaq, thetaq = y[..., -2], y[..., -1]
    b = _bg(p, a)
    cs2, pb43, w_Q, rho_Q, H, ca2_Q, opac, xe = b["cs2"], b["pb43"], b["w_Q"], b["rho_Q"], b["H"], b["ca2_Q"], b["opac"], b["xe"]
    cs2_Q, wa = p["cs2_DE"], p["w_DE_a"]
    G = p["grhom"] * p["OmegaDE"]
    k2 = k ** 2
    loga = np.log(a)
    # coefficient derivatives
    sc = p["cs2a_of_loga_spline"].evaluate(loga)
    dcs2 = _spl_d(p["cs2a_of_loga_spline"], loga) / a ** 2 - sc / a ** 2
    dxe = _spl_d(p["xe_of_loga_spline"], loga) / a
    dpb43 = -pb43 / a
    drhoQ = rho_Q * (-3 * (1 + p["w_DE_0"] + wa) / a + 3 * wa)
    rhonu = np.exp(p["logrhonu_of_loga_spline"].evaluate(loga))
    drhonu_bg = rhonu * _spl_d(p["logrhonu_of_loga_spline"], loga) / a
    dgq = G * (drhoQ * a ** 2 + 2 * a * rho_Q)                 # d/da of grhom OmegaDE rho_Q a^2
    dgrho_bg = (-p["grhom"] * p["Omegam"] / a ** 2 - 2 * (p["grhog"] + p["grhor"] * (p["Neff"] + p["Nmnu"] * rhonu)) / a ** 3
                + p["grhor"] * p["Nmnu"] * drhonu_bg / a ** 2 + dgq)
    dH = dgrho_bg / (6 * H)
    Dq = (1 + w_Q) + 1e-6
    dca2 = -wa + wa / (3 * Dq) + wa ** 2 * a / (3 * Dq ** 2)
    akthom = AKTHOM_RHS * (1.0 - p["YHe"]) * p["Omegab"] * p["H0"] ** 2
    dopac = akthom * (dxe / a ** 2 - 2 * xe / a ** 3)
    aq = a[..., None] * p["amnu"] / d.q
    v = 1 / np.sqrt(1 + aq ** 2)
    dv = -v ** 3 * aq * p["amnu"] / d.q
    psi0, psi1, psi2 = y[..., iq0:iq0 + nq], y[..., iq0 + nq:iq0 + 2 * nq], y[..., iq0 + 2 * nq:iq0 + 3 * nq]
    psi3 = y[..., iq0 + 3 * nq:iq0 + 4 * nq]
    drhonu = np.sum(d.w * psi0 / v, -1)
    dpnu = np.sum(d.w * psi0 * v, -1) / 3
    fnu = np.sum(d.w * psi1, -1)
    ddrhonu = np.sum(d.w * psi0 * (-dv / v ** 2), -1)
    ddpnu = np.sum(d.w * psi0 * dv, -1) / 3
    rpt = (1 + w_Q) * rho_Q * G * thetaq * a ** 2
    drpt = G * thetaq * (-wa * rho_Q * a ** 2 + (1 + w_Q) * (drhoQ * a ** 2 + 2 * a * rho_Q))
    # metric sources
    m1 = p["grhom"] * (Omegac * deltac + p["Omegab"] * deltab)
    m2 = p["grhog"] * deltag + p["grhor"] * (p["Neff"] * deltar + p["Nmnu"] * drhonu)
    dgrho = m1 / a + m2 / a ** 2 + G * deltaq * rho_Q * a ** 2
    ddgrho = -m1 / a ** 2 - 2 * m2 / a ** 3 + p["grhor"] * p["Nmnu"] * ddrhonu / a ** 2 + deltaq * dgq
    p1 = p["grhog"] * deltag + p["grhor"] * p["Neff"] * deltar
    dgpres = (p1 / a ** 2 / 3.0 + p["grhor"] * p["Nmnu"] * dpnu / a ** 2
              + (cs2_Q * G * deltaq * rho_Q * a ** 2 + (cs2_Q - ca2_Q) * (3 * H * rpt / k2)))
    ddgpres = (-2 * p1 / a ** 3 / 3.0 + p["grhor"] * p["Nmnu"] * (ddpnu / a ** 2 - 2 * dpnu / a ** 3)
               + cs2_Q * deltaq * dgq - dca2 * (3 * H * rpt / k2) + (cs2_Q - ca2_Q) * 3 * (dH * rpt + H * drpt) / k2)
    t1 = p["grhom"] * (Omegac * thetac + p["Omegab"] * thetab)
    t2 = 4.0 / 3.0 * (p["grhog"] * thetag + p["Neff"] * p["grhor"] * thetar) + p["Nmnu"] * p["grhor"] * k * fnu
    ddgtheta = -t1 / a ** 2 - 2 * t2 / a ** 3 + drpt
    hprime = (2.0 * k2 * eta + dgrho) / H
    dhp = ddgrho / H - hprime * dH / H
    dep = 0.5 * ddgtheta / k2
    dal = (dhp + 6.0 * dep) / 2.0 / k2

    f = np.zeros_like(y)
    f[..., 0] = dH * a + H
    f[..., 1] = -(ddgrho + 3.0 * ddgpres) * a - (dgrho + 3.0 * dgpres)
    f[..., 2] = dep
    f[..., 3] = -0.5 * dhp
    f[..., 4] = -dH * thetac
    f[..., 5] = -0.5 * dhp
    f[..., 6] = -dH * thetab + k2 * dcs2 * deltab + (dpb43 * opac + pb43 * dopac) * (thetag - thetab)
    polter = y[..., ig + 2] + y[..., igp] + y[..., igp + 2]
    dop = dopac[..., None]
    f[..., ig] = -2.0 / 3.0 * dhp
    f[..., ig + 1] = -dopac * (thetag - thetab)
    f[..., ig + 3:ig + Lg + 1] = -dop * y[..., ig + 3:ig + Lg + 1]
    f[..., ig + 2] = 8.0 / 15.0 * k2 * dal - dopac * (y[..., ig + 2] - 0.1 * polter)
    f[..., igp:igp + Lp + 1] = -dop * y[..., igp:igp + Lp + 1]
    f[..., igp] += dopac * polter / 2
    f[..., igp + 2] += dopac * polter / 10
    f[..., ir] = -2.0 / 3.0 * dhp
    f[..., ir + 2] = 8.0 / 15.0 * k2 * dal
    dl = d.dlfdlq
    kk = k[..., None]
    f[..., iq0:iq0 + nq] = -kk * dv * psi1 + dhp[..., None] * dl / 6.0
    f[..., iq0 + nq:iq0 + 2 * nq] = kk * dv * (psi0 - 2.0 * psi2) / 3.0
    f[..., iq0 + 2 * nq:iq0 + 3 * nq] = kk * dv * (2 * psi1 - 3 * psi3) / 5.0 - (dhp / 15 + 2 / 5 * dep)[..., None] * dl
    for l in range(3, Ln):
        lo = y[..., iq0 + (l - 1) * nq:iq0 + l * nq]
        hi = y[..., iq0 + (l + 1) * nq:iq0 + (l + 2) * nq]
        f[..., iq0 + l * nq:iq0 + (l + 1) * nq] = kk * dv / (2 * l + 1) * (l * lo - (l + 1) * hi)
    f[..., iq0 + Ln * nq:iq0 + (Ln + 1) * nq] = kk * dv * y[..., iq0 + (Ln - 1) * nq:iq0 + Ln * nq]
    f[..., -2] = (wa * (thetaq + 0.5 * hprime) - (1 + w_Q) * 0.5 * dhp - 3 * (wa * H + (cs2_Q - w_Q) * dH) * deltaq
                  - 9 * thetaq / k2 * (-wa * (cs2_Q - ca2_Q) * H ** 2 - (1 + w_Q) * dca2 * H ** 2
                                       + (1 + w_Q) * (cs2_Q - ca2_Q) * 2 * H * dH))
    f[..., -1] = -(1 - 3 * cs2_Q) * dH * thetaq + cs2_Q * k2 * deltaq * wa / (1 + w_Q) ** 2
    return f


def jacobian_bruteforce(tau, y, p, k, d: Dims):
    """Dense df/dy exactly as forward-mode AD sees it: the RHS is linear in y[1:], so column
    j>=1 is f(a, e_j) - f(a, 0); column 0 (the scale factor) by the complex step."""
    n = d.n
    M = y.shape[0]
    J = np.zeros((M, n, n))
    base = np.zeros_like(y)
    base[:, 0] = y[:, 0]
    f0 = rhs(tau, base, p, k, d)
    for j in range(1, n):
        yy = base.copy()
        yy[:, j] = 1.0
        J[:, :, j] = rhs(tau, yy, p, k, d) - f0
    h = 1e-30
    yc = y.astype(np.complex128)
    yc[:, 0] += 1j * h * y[:, 0]
    J[:, :, 0] = rhs(tau, yc, p, k, d).imag / (h * y[:, 0])[:, None]
    return J


def jacobian(tau, y, p, k, d: Dims):
    """Dense J[M, n, n] = df/dy assembled analytically column block by column block; the
    a-column uses the complex step on the restated RHS (exact to round-off)."""
    n, nq, iq0 = d.n, d.nq, d.iq0
    ig, igp, ir = d.ig, d.igp, d.ir
    Lg, Lp, Lr, Ln = d.lmaxg, d.lmaxgp, d.lmaxr, d.lmaxnu
    M = y.shape[0]
    cplx = np.iscomplexobj(y)            # only under the tangent oracle's complex step (oracle/discoeb_tangent.py)
    dtype = np.complex128 if cplx else np.float64
    J = np.zeros((M, n, n), dtype=dtype)
    a = y[:, 0]
    b = _bg(p, a)
    cs2, pb43, w_Q, rho_Q, H, ca2_Q, opac = b["cs2"], b["pb43"], b["w_Q"], b["rho_Q"], b["H"], b["ca2_Q"], b["opac"]
    cs2_Q = p["cs2_DE"]
    Omegac = p["Omegam"] - p["Omegab"]
    k2 = k ** 2
    aq = a[:, None] * p["amnu"] / d.q
    v = 1 / np.sqrt(1 + aq ** 2)
    dl = d.dlfdlq

    # gradients of the metric sources (rows of length n)
    g_rho = np.zeros((M, n), dtype=dtype)
    g_pres = np.zeros((M, n), dtype=dtype)
    g_th = np.zeros((M, n), dtype=dtype)
    g_rho[:, 3] = p["grhom"] * Omegac / a
    g_rho[:, 5] = p["grhom"] * p["Omegab"] / a
    g_rho[:, 7] = p["grhog"] / a ** 2
    g_rho[:, ir] = p["grhor"] * p["Neff"] / a ** 2
    g_rho[:, iq0:iq0 + nq] = (p["grhor"] * p["Nmnu"] / a ** 2)[:, None] * d.w / v
    g_rho[:, -2] = p["grhom"] * p["OmegaDE"] * rho_Q * a ** 2
    rptc = (1 + w_Q) * rho_Q * p["grhom"] * p["OmegaDE"] * a ** 2       # d(rho+p)theta_Q / d thetaq
    g_pres[:, 7] = p["grhog"] / a ** 2 / 3.0
    g_pres[:, ir] = p["grhor"] * p["Neff"] / a ** 2 / 3.0
    g_pres[:, iq0:iq0 + nq] = (p["grhor"] * p["Nmnu"] / a ** 2)[:, None] * d.w * v / 3
    g_pres[:, -2] = cs2_Q * p["grhom"] * p["OmegaDE"] * rho_Q * a ** 2
    g_pres[:, -1] = (cs2_Q - ca2_Q) * 3 * H * rptc / k2
    g_th[:, 4] = p["grhom"] * Omegac / a
    g_th[:, 6] = p["grhom"] * p["Omegab"] / a
    g_th[:, 8] = 4.0 / 3.0 * p["grhog"] / a ** 2
    g_th[:, ir + 1] = 4.0 / 3.0 * p["Neff"] * p["grhor"] / a ** 2
    g_th[:, iq0 + nq:iq0 + 2 * nq] = (p["Nmnu"] * p["grhor"] * k / a ** 2)[:, None] * d.w
    g_th[:, -1] = rptc
    g_h = g_rho / H[:, None]
    g_h[:, 2] += 2.0 * k2 / H
    g_e = 0.5 * g_th / k2[:, None]
    g_al = (g_h + 6.0 * g_e) / 2.0 / k2[:, None]

    J[:, 1, :] = -(g_rho + 3.0 * g_pres) * a[:, None]
    J[:, 2, :] = g_e
    J[:, 3, :] = -0.5 * g_h
    J[:, 3, 4] += -1.0
    J[:, 4, 4] = -H
    J[:, 5, :] = -0.5 * g_h
    J[:, 5, 6] += -1.0
    J[:, 6, 6] += -H - pb43 * opac
    J[:, 6, 5] += k2 * cs2
    J[:, 6, 8] += pb43 * opac
    # photons
    J[:, ig, :] += -2.0 / 3.0 * g_h
    J[:, ig, 8] += -4.0 / 3.0
    J[:, ig + 1, 7] += 0.25 * k2
    J[:, ig + 1, 9] += -0.5 * k2
    J[:, ig + 1, 8] += -opac
    J[:, ig + 1, 6] += opac
    J[:, ig + 2, :] += 8.0 / 15.0 * k2[:, None] * g_al
    J[:, ig + 2, 8] += 8.0 / 15.0
    J[:, ig + 2, ig + 3] += -0.6 * k
    J[:, ig + 2, ig + 2] += -opac * 0.9
    J[:, ig + 2, igp] += 0.1 * opac
    J[:, ig + 2, igp + 2] += 0.1 * opac
    for l in range(3, Lg):
        J[:, ig + l, ig + l - 1] += k * l / (2 * l + 1)
        J[:, ig + l, ig + l + 1] += -k * (l + 1) / (2 * l + 1)
        J[:, ig + l, ig + l] += -opac
    J[:, ig + Lg, ig + Lg - 1] += k
    J[:, ig + Lg, ig + Lg] += -(Lg + 1) / tau - opac
    for l in range(0, Lp):
        if l > 0:
            J[:, igp + l, igp + l - 1] += k * l / (2 * l + 1)
        J[:, igp + l, igp + l + 1] += -k * (l + 1) / (2 * l + 1)
        J[:, igp + l, igp + l] += -opac
    for row, fac in ((igp, 0.5), (igp + 2, 0.1)):
        for col in (ig + 2, igp, igp + 2):
            J[:, row, col] += opac * fac
    J[:, igp + Lp, igp + Lp - 1] += k
    J[:, igp + Lp, igp + Lp] += -(Lp + 1) / tau - opac
    # massless nu
    J[:, ir, :] += -2.0 / 3.0 * g_h
    J[:, ir, ir + 1] += -4.0 / 3.0
    J[:, ir + 1, ir] += 0.25 * k2
    J[:, ir + 1, ir + 2] += -0.5 * k2
    J[:, ir + 2, :] += 8.0 / 15.0 * k2[:, None] * g_al
    J[:, ir + 2, ir + 1] += 8.0 / 15.0
    J[:, ir + 2, ir + 3] += -0.6 * k
    for l in range(3, Lr):
        J[:, ir + l, ir + l - 1] += k * l / (2 * l + 1)
        J[:, ir + l, ir + l + 1] += -k * (l + 1) / (2 * l + 1)
    J[:, ir + Lr, ir + Lr - 1] += k
    J[:, ir + Lr, ir + Lr] += -(Lr + 1) / tau
    # massive nu
    for i in range(nq):
        kv = k * v[:, i]
        r0, r1, r2 = iq0 + i, iq0 + nq + i, iq0 + 2 * nq + i
        J[:, r0, :] += g_h * dl[i] / 6.0
        J[:, r0, r1] += -kv
        J[:, r1, r0] += kv / 3.0
        J[:, r1, r2] += -2.0 * kv / 3.0
        J[:, r2, :] += -(g_h / 15 + 2 / 5 * g_e) * dl[i]
        J[:, r2, r1] += 2 * kv / 5.0
        J[:, r2, iq0 + 3 * nq + i] += -3 * kv / 5.0
        for l in range(3, Ln):
            r = iq0 + l * nq + i
            J[:, r, r - nq] += kv * l / (2 * l + 1)
            J[:, r, r + nq] += -kv * (l + 1) / (2 * l + 1)
        r = iq0 + Ln * nq + i
        J[:, r, r - nq] += kv
        J[:, r, r] += -(Ln + 1) / tau
    # dark energy
    J[:, -2, :] += -(1 + w_Q)[:, None] * 0.5 * g_h
    J[:, -2, -1] += -(1 + w_Q) - 9 * (1 + w_Q) * (cs2_Q - ca2_Q) * H ** 2 / k2
    J[:, -2, -2] += -3 * (cs2_Q - w_Q) * H
    J[:, -1, -1] += -(1 - 3 * cs2_Q) * H
    J[:, -1, -2] += cs2_Q / (1 + w_Q) * k2
    # scale-factor column by complex step (every coefficient depends on a); when the state is already
    # complex (an outer complex step is in flight) the hand-derived column is used instead
    if cplx:
        J[:, :, 0] = rhs_da(tau, y, p, k, d)
        return J
    h = 1e-30
    yc = y.astype(np.complex128)
    yc[:, 0] += 1j * h * a
    J[:, :, 0] = rhs(tau, yc, p, k, d).imag / (h * a)[:, None]
    return J


def dfdt(tau, y, d: Dims):
    """Explicit tau-derivative: only the truncation rows carry (lmax+1)/tau (App. B)."""
    dT = np.zeros_like(y)
    t2 = tau ** 2
    for base, L in ((d.ig, d.lmaxg), (d.igp, d.lmaxgp), (d.ir, d.lmaxr)):
        dT[:, base + L] = (L + 1) / t2 * y[:, base + L]
    nq, Ln = d.nq, d.lmaxnu
    sl = slice(d.iq0 + Ln * nq, d.iq0 + (Ln + 1) * nq)
    dT[:, sl] = (Ln + 1) / t2[:, None] * y[:, sl]
    return dT


def rodas5_step(t0, t1, y0, p, k, d: Dims, jac=jacobian):
    """One attempted Rodas5Transformed step for a batch of modes (ode_integrators_stiff.py:718-834)."""
    dt = t1 - t0
    n = d.n
    Jm = jac(t0, y0, p, k, d)
    W = np.eye(n)[None] / (dt * GAMMA)[:, None, None] - Jm
    lu, piv = sla.lu_factor(W, check_finite=False)

    def solve(r):
        return sla.lu_solve((lu, piv), r[..., None], check_finite=False)[..., 0]

    dT = dfdt(t0, y0, d)
    dtc = dt[:, None]
    ks = {}
    ks[1] = solve(rhs(t0, y0, p, k, d) + dtc * D[1] * dT)
    u = None
    for i in range(2, 9):
        if i <= 6:
            u = y0 + sum(A[(i, j)] * ks[j] for j in range(1, i))
        else:
            u = u + ks[i - 1]
        ti = t0 + CT[i] * dt if i <= 5 else t0 + dt
        r = rhs(ti, u, p, k, d)
        if i <= 5:
            r = r + dtc * D[i] * dT
        r = r + sum((C[(i, j)] / dtc) * ks[j] for j in range(1, i))
        ks[i] = solve(r)
    y1 = u + ks[8]
    return y1, ks[8]


def scaled_error_norm(y0, y1, err, k, rtol, atol):
    """diffrax PIDController scaling + the reference's filtered RMS norm (perturbations.py:701-711, :759)."""
    nan = np.isnan(y1).any(-1, keepdims=True)
    y1c = np.where(nan, y0, y1)
    sc = err / (atol + np.maximum(np.abs(y0), np.abs(y1c)) * rtol)
    idx = np.array([0, 2, 3, 5, 6, 7])
    wts = np.stack([np.ones_like(k), k ** 2, np.ones_like(k), np.ones_like(k), 1 / k ** 2, np.ones_like(k)], -1)
    x = sc[:, idx] * wts
    return np.sqrt(np.mean(x * x, -1))


def integrate_modes(t0, t1, y0, ts, p, k, d: Dims, rtol, atol, pcoeff=0.25, icoeff=0.8, dcoeff=0.0,
                    factormax=20.0, factormin=0.3, max_steps=2048, safety=0.9, order=5.0, jac=jacobian,
                    trace=None):
    """diffrax.diffeqsolve semantics for M independent modes advanced in lock-step
    with masks -- what jax.vmap of the reference's per-mode solve does.

    t0[M] start times, t1 scalar end time, ts[nout] ascending save times.  Returns
    (ys[M, nout, n], status[M] (0 ok / 1 max_steps / 2 non-finite), nsteps[M] attempted, naccept[M]).
    """
    M, n = y0.shape
    nout = len(ts)
    ys = np.full((M, nout, n), np.nan)
    tprev = t0.astype(np.float64).copy()
    dt0 = np.minimum(t0 / 4, 0.5 * (t1 - t0))            # perturbations.py:756
    tnext = tprev + dt0
    tnext = np.where(tnext > t1 - 1e-10, t1, tnext)      # _clip_to_end(keep_step=True) at init
    y = y0.copy()
    inv_prev = np.ones(M)
    inv_pprev = np.ones(M)
    save_idx = np.zeros(M, dtype=np.int64)
    nsteps = np.zeros(M, dtype=np.int64)
    nacc = np.zeros(M, dtype=np.int64)
    status = np.zeros(M, dtype=np.int64)
    c1 = (icoeff + pcoeff + dcoeff) / order
    c2 = -(pcoeff + 2 * dcoeff) / order
    c3 = dcoeff / order
    while True:
        act = np.nonzero((tprev < t1) & (status == 0) & (nsteps < max_steps))[0]
        if act.size == 0:
            break
        tp, tn, ya, ka = tprev[act], tnext[act], y[act], k[act]
        with np.errstate(all="ignore"):
            y1, err = rodas5_step(tp, tn, ya, p, ka, d, jac=jac)
            err = np.where(np.isnan(err), np.inf, err)
            E = scaled_error_norm(ya, y1, err, ka, rtol, atol)
            keep = E < 1
            inv = 1.0 / E
            f1 = inv ** c1 if c1 != 0 else 1.0
            f2 = inv_prev[act] ** c2 if c2 != 0 else 1.0
            f3 = inv_pprev[act] ** c3 if c3 != 0 else 1.0
            fac = np.clip(safety * f1 * f2 * f3, np.where(keep, 1.0, factormin), factormax)
            dtn = (tn - tp) * fac
            inv = np.where((inv == 0) | np.isinf(inv), 1.0, inv)
        nsteps[act] += 1
        nacc[act] += keep
        if trace is not None:
            trace.append((act.copy(), tp.copy(), tn.copy(), E.copy(), keep.copy()))
        # SaveAt(ts): linear interpolation inside every accepted step (LocalLinearInterpolation)
        for m in np.nonzero(keep)[0]:
            g = act[m]
            while save_idx[g] < nout and ts[save_idx[g]] <= tn[m]:
                tt = ts[save_idx[g]]
                coeff = 0.0 if tn[m] == tp[m] else (tt - tp[m]) / (tn[m] - tp[m])
                ys[g, save_idx[g]] = ya[m] + coeff * (y1[m] - ya[m])
                save_idx[g] += 1
        y[act] = np.where(keep[:, None], y1, ya)
        tpn = np.minimum(np.where(keep, tn, tp), t1)
        tnn = tpn + dtn
        clip = tnn > t1 - 1e-10
        tnn = np.where(clip, np.where(keep, t1, tpn + 0.5 * (t1 - tpn)), tnn)
        tprev[act], tnext[act] = tpn, tnn
        inv_pprev[act] = np.where(keep, inv_prev[act], inv_pprev[act])
        inv_prev[act] = np.where(keep, inv, inv_prev[act])
        bad = ~np.isfinite(tnn) | (~np.isfinite(y[act])).any(-1)
        status[act[bad]] = 2
    status[(status == 0) & (tprev < t1)] = 1
    return ys, status, nsteps, nacc


def _bisect(func, xl, xr, numit):
    """util.py:365-396 (keeps [mid, right] when f(mid) f(left) > 0)."""
    xl = np.array(xl, dtype=np.result_type(xl, np.float64), copy=True)      # (complex under the tangent oracle)
    xr = np.array(xr, dtype=np.result_type(xr, np.float64), copy=True)
    for _ in range(numit):
        xm = 0.5 * (xl + xr)
        c = np.real(func(xm)) * np.real(func(xl)) > 0
        xl, xr = np.where(c, xm, xl), np.where(c, xr, xm)
    return 0.5 * (xl + xr)


def determine_starting_time(p, k):
    """perturbations.py:630-681."""
    k = np.asarray(k, dtype=np.result_type(np.asarray(k).dtype, np.float64))      # (complex under the tangent oracle)
    tau0 = p["taumin"]
    tau1 = p["tau_of_a_spline"].evaluate(0.1)
    tau_k = 1.0 / k
    akthom = AKTHOM_START * (1.0 - p["YHe"]) * p["Omegab"] * p["H0"] ** 2

    def cond_small_k(lt):
        tau = np.exp(lt)
        xe = p["xe_of_tau_spline"].evaluate(tau)
        a = p["a_of_tau_spline"].evaluate(tau)
        opac = xe * akthom / a ** 2
        H = aprimeoa(p, a)
        return (1.0 / opac) / (1.0 / H) / 0.0004 - 1.0

    def cond_large_k(lt):
        a = p["a_of_tau_spline"].evaluate(np.exp(lt))
        return (1.0 / aprimeoa(p, a)) / tau_k / 0.07 - 1.0

    xl = np.full(k.shape, np.log(tau0))
    xr = np.full(k.shape, np.log(tau1))
    lt_large = _bisect(cond_large_k, xl, xr, 7)
    lt_small = _bisect(cond_small_k, xl, xr, 7)
    return np.exp(np.where(np.real(lt_small) <= np.real(lt_large), lt_small, lt_large))


def adiabatic_ics(tau, p, k, d: Dims):
    """perturbations.py:526-627 for a batch of modes."""
    M = k.shape[0]
    a = p["a_of_tau_spline"].evaluate(tau)
    y = np.zeros((M, d.n), dtype=np.result_type(a, np.float64))     # complex only under the tangent oracle's complex step
    rhonu_s = np.exp(p["logrhonu_of_loga_spline"].evaluate(np.log(a)))
    rhom = p["grhom"] * p["Omegam"] / a ** 3
    rhor = (p["grhog"] + p["grhor"] * (p["Neff"] + p["Nmnu"] * rhonu_s)) / a ** 4
    rhonu = p["grhor"] * (p["Neff"] + p["Nmnu"] * rhonu_s) / a ** 4
    fracb = p["Omegab"] / p["Omegam"]
    fracnu = rhonu / rhor
    om = a * rhom / np.sqrt(rhor)
    ci = -1.0
    s2 = 1.0
    kt = k * tau
    deltag = -kt ** 2 / 3 * (1 - om * tau / 5) * ci * s2
    thetag = -kt ** 3 / tau / 36 * (1 - 3 * (1 + 5 * fracb - fracnu) / 20 / (1 - fracnu) * om * tau) * ci * s2
    deltar = deltag
    thetar = -kt ** 4 / tau / 36 / (4 * fracnu + 15) * (4 * fracnu + 11 + 12 - 3 * (8 * fracnu * fracnu + 50 * fracnu + 275) / 20 / (2 * fracnu + 15) * tau * om) * ci
    shearr = kt ** 2 / (45 + 12 * fracnu) * (3 * s2 - 1) * (1 + (4 * fracnu - 5) / 4 / (2 * fracnu + 15) * tau * om) * ci
    cs2_Q = p["cs2_DE"]
    w_Q = p["w_DE_0"] + p["w_DE_a"] * (1.0 - a)
    deltaq = kt ** 2 / 4 * (1 + w_Q) * (4 - 3 * cs2_Q) / (4 - 6 * w_Q + 3 * cs2_Q) * ci * s2
    thetaq = kt ** 4 / tau / 4 * cs2_Q / (4 - 6 * w_Q + 3 * cs2_Q) * ci * s2
    eta = ci * (1 - kt ** 2 / 12 / (15 + 4 * fracnu) * (5 + 4 * s2 * fracnu - (16 * fracnu * fracnu + 280 * fracnu + 325) / 10 / (2 * fracnu + 15) * tau * om))
    y[:, 0] = a
    y[:, 2] = eta
    y[:, 3] = 0.75 * deltag
    y[:, 5] = 0.75 * deltag
    y[:, 6] = thetag
    y[:, 7] = deltag
    y[:, 8] = thetag
    y[:, d.ir] = deltar
    y[:, d.ir + 1] = thetar
    y[:, d.ir + 2] = shearr * 2.0
    nq, iq0 = d.nq, d.iq0
    aq = a[:, None] * p["amnu"] / d.q
    v = 1 / np.sqrt(1 + aq ** 2)
    dl = d.dlfdlq
    y[:, iq0:iq0 + nq] = -0.25 * dl * deltar[:, None]
    y[:, iq0 + nq:iq0 + 2 * nq] = -dl * thetar[:, None] / v / k[:, None] / 3.0
    y[:, iq0 + 2 * nq:iq0 + 3 * nq] = -0.5 * dl * shearr[:, None]
    y[:, -2] = deltaq
    y[:, -1] = thetaq
    return y


def convert_to_output(y, p, k, d: Dims):
    """State -> 20 gauge-fixed output fields; y[..., n], k broadcastable (perturbations.py:374-523)."""
    nq, iq0 = d.nq, d.iq0
    Omegac = p["Omegam"] - p["Omegab"]
    a, eta = y[..., 0], y[..., 2]
    deltac, thetac, deltab, thetab, deltag, thetag = (y[..., i] for i in range(3, 9))
    deltar, thetar = y[..., d.ir], y[..., d.ir + 1]
    la = np.log(a)
    rhonu = np.exp(p["logrhonu_of_loga_spline"].evaluate(la))
    pnu = np.exp(p["logpnu_of_loga_spline"].evaluate(la))
    aq = a[..., None] * p["amnu"] / d.q
    v = 1 / np.sqrt(1 + aq ** 2)
    drhonu = np.sum(d.w * y[..., iq0:iq0 + nq] / v, -1)
    fnu = np.sum(d.w * y[..., iq0 + nq:iq0 + 2 * nq], -1)
    deltanu = drhonu / rhonu
    thetanu = k * fnu / (rhonu + pnu)
    deltaq, thetaq = y[..., -2], y[..., -1]
    w_Q = p["w_DE_0"] + p["w_DE_a"] * (1.0 - a)
    rho_Q = a ** (-3 * (1 + p["w_DE_0"] + p["w_DE_a"])) * np.exp(3 * (a - 1) * p["w_DE_a"])
    rpt_Q = (1 + w_Q) * rho_Q * p["grhom"] * p["OmegaDE"] * thetaq * a ** 2
    grho = (p["grhom"] * p["Omegam"] / a + (p["grhog"] + p["grhor"] * (p["Neff"] + p["Nmnu"] * rhonu)) / a ** 2
            + p["grhom"] * p["OmegaDE"] * rho_Q * a ** 2 + p["grhom"] * p["Omegak"])
    H = np.sqrt(grho / 3.0)
    mat = p["grhom"] * (Omegac * deltac + p["Omegab"] * deltab) / a
    matth = p["grhom"] * (Omegac * thetac + p["Omegab"] * thetab) / a
    dgrho = (mat + (p["grhog"] * deltag + p["grhor"] * (p["Neff"] * deltar + p["Nmnu"] * drhonu)) / a ** 2
             + p["grhom"] * p["OmegaDE"] * deltaq * rho_Q * a ** 2)
    dgtheta = (matth + 4.0 / 3.0 * (p["grhog"] * thetag + p["Neff"] * p["grhor"] * thetar) / a ** 2
               + p["Nmnu"] * p["grhor"] * k * fnu / a ** 2 + rpt_Q)
    k2 = k ** 2
    hprime = (2.0 * k2 * eta + dgrho) / H
    etaprime = 0.5 * dgtheta / k2
    alpha = (hprime + 6.0 * etaprime) / 2.0 / k2
    deltam = (mat + (p["grhor"] * p["Nmnu"] * drhonu) / a ** 2) / (p["grhom"] * p["Omegam"] / a + (p["grhor"] * p["Nmnu"] * rhonu) / a ** 2)
    thetam = (matth + p["Nmnu"] * p["grhor"] * k * fnu / a ** 2) / (3.0 * (p["grhom"] * p["Omegam"] / a + p["grhor"] * p["Nmnu"] * rhonu / a ** 2))
    deltabc = mat / (p["grhom"] * p["Omegam"] / a)
    thetabc = matth / (3.0 * (p["grhom"] * p["Omegam"] / a) / a ** 2)
    thetam = thetam + alpha * k2
    thetabc = thetabc + alpha * k2
    return np.stack([eta, etaprime, hprime, alpha, deltam, thetam / H, deltabc, thetabc / H,
                     deltac, thetac / H, deltab, thetab / H, deltag, thetag / H, deltar, thetar / H,
                     deltanu, thetanu / H, deltaq, thetaq / H], -1)


def evolve_perturbations(*, param, aexp_out, kmin, kmax, num_k, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8,
                         nqmax=3, rtol=1e-4, atol=1e-4, pcoeff=0.25, icoeff=0.80, dcoeff=0.0, factormax=20.0,
                         factormin=0.3, max_steps=2048, return_full=False, dologk=True, kmodes=None,
                         chunk=64, return_info=False):
    """perturbations.py:926-997 (+ evolve_one_mode :728-781).  ``kmodes`` overrides the k grid
    (oracle-only convenience for sub-sampled parity tests)."""
    if kmodes is None:
        kmodes = np.geomspace(kmin, kmax, num_k) if dologk else np.linspace(kmin, kmax, num_k)
    kmodes = np.asarray(kmodes, dtype=np.float64)
    aexp_out = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    tau_out = param["tau_of_a_spline"].evaluate(aexp_out)
    tau_max = np.max(tau_out)
    d = Dims(lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax)
    M = kmodes.shape[0]
    nout = tau_out.shape[0]
    yfull = np.zeros((M, nout, d.n))
    status = np.zeros(M, dtype=np.int64)
    nsteps = np.zeros(M, dtype=np.int64)
    nacc = np.zeros(M, dtype=np.int64)
    tau_start = 0.99 * np.minimum(np.min(tau_out), determine_starting_time(param, kmodes))
    for s in range(0, M, chunk):
        sl = slice(s, min(M, s + chunk))
        y0 = adiabatic_ics(tau_start[sl], param, kmodes[sl], d)
        yfull[sl], status[sl], nsteps[sl], nacc[sl] = integrate_modes(
            tau_start[sl], tau_max, y0, tau_out, param, kmodes[sl], d, rtol, atol, pcoeff, icoeff, dcoeff,
            factormax, factormin, max_steps)
    if np.any(status != 0):
        raise RuntimeError(f"oracle: {np.count_nonzero(status)} modes failed (status {np.unique(status)})")
    yout = yfull if return_full else convert_to_output(yfull, param, kmodes[:, None], d)
    param["lmaxg"], param["lmaxgp"], param["lmaxr"], param["lmaxnu"] = lmaxg, lmaxgp, lmaxr, lmaxnu
    param["nqmax"], param["nout"], param["tau_out"] = nqmax, nout, tau_out
    if return_info:
        return yout, kmodes, param, dict(nsteps=nsteps, naccept=nacc, tau_start=tau_start, status=status)
    return yout, kmodes, param


def get_power(*, k, y, idx, param):
    """perturbations.py:1101-1123."""
    return 2 * np.pi ** 2 * param["A_s"] * (k / param["k_p"]) ** (param["n_s"] - 1) * k ** (-3) * y[..., idx] ** 2


# --------------------------------------------------------------------------------------------------
# batched variant: one shared step size per batch of modes
# (perturbations.py:786-922 evolve_modes_batched, ode_integrators_stiff.py:846-1010 Rodas5Batched)
# --------------------------------------------------------------------------------------------------
def integrate_batch(t0, t1, y0, ts, p, k, d: Dims, rtol, atol, pcoeff=0.25, icoeff=0.8, dcoeff=0.0, factormax=20.0,
                    factormin=0.3, max_steps=2048, safety=0.9, order=5.0, trace=None):
    """diffrax.diffeqsolve of ONE batch: state y[B, n] is a single pytree, so there is one t, one dt, one
    accept/reject decision and one error norm -- the RMS over all B x 6 filtered, weighted components
    (rms_norm_filtered_batched, perturbations.py:689-693); a NaN anywhere in the batch's candidate replaces the whole
    candidate by y0 in the error scale (diffrax's ``jnp.isnan(y1).any()`` acts on the whole array).
    Returns (ys[B, nout, n], status, nsteps, naccept) -- scalars shared by the batch."""
    B, n = y0.shape
    nout = len(ts)
    ys = np.full((B, nout, n), np.nan)
    tprev = float(t0)
    tnext = tprev + min(tprev / 4, 0.5 * (t1 - tprev))
    if tnext > t1 - 1e-10:
        tnext = t1
    y = y0.copy()
    inv_prev = inv_pprev = 1.0
    save_idx = nsteps = nacc = status = 0
    c1 = (icoeff + pcoeff + dcoeff) / order
    c2 = -(pcoeff + 2 * dcoeff) / order
    c3 = dcoeff / order
    idx = np.array([0, 2, 3, 5, 6, 7])
    wts = np.stack([np.ones_like(k), k ** 2, np.ones_like(k), np.ones_like(k), 1 / k ** 2, np.ones_like(k)], -1)
    while tprev < t1 and status == 0 and nsteps < max_steps:
        tp = np.full(B, tprev)
        tn = np.full(B, tnext)
        with np.errstate(all="ignore"):
            y1, err = rodas5_step(tp, tn, y, p, k, d)
            err = np.where(np.isnan(err), np.inf, err)
            y1c = y if np.isnan(y1).any() else y1
            sc = err / (atol + np.maximum(np.abs(y), np.abs(y1c)) * rtol)
            x = sc[:, idx] * wts
            E = float(np.sqrt(np.mean(x * x)))
            keep = E < 1
            inv = 1.0 / E if E != 0 else np.inf
            f1 = inv ** c1 if c1 != 0 else 1.0
            f2 = inv_prev ** c2 if c2 != 0 else 1.0
            f3 = inv_pprev ** c3 if c3 != 0 else 1.0
            fac = float(np.clip(safety * f1 * f2 * f3, 1.0 if keep else factormin, factormax))
            dtn = (tnext - tprev) * fac
            if inv == 0 or np.isinf(inv):
                inv = 1.0
        nsteps += 1
        nacc += int(keep)
        if trace is not None:
            trace.append((tprev, tnext, E, keep))
        if keep:
            while save_idx < nout and ts[save_idx] <= tnext:
                tt = ts[save_idx]
                coeff = 0.0 if tnext == tprev else (tt - tprev) / (tnext - tprev)
                ys[:, save_idx] = y + coeff * (y1 - y)
                save_idx += 1
            y = y1
            inv_pprev, inv_prev = inv_prev, inv
            tprev = min(tnext, t1)
        tn_ = tprev + dtn
        if tn_ > t1 - 1e-10:
            tn_ = t1 if keep else tprev + 0.5 * (t1 - tprev)
        tnext = tn_
        if not np.isfinite(tnext) or not np.all(np.isfinite(y)):
            status = 2
    if status == 0 and tprev < t1:
        status = 1
    return ys, status, nsteps, nacc


def evolve_perturbations_batched(*, param, aexp_out, kmin, kmax, num_k, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3,
                                 rtol=1e-4, atol=1e-4, pcoeff=0.25, icoeff=0.80, dcoeff=0.0, factormax=20.0, factormin=0.3,
                                 max_steps=2048, batch_size=16, kmodes=None, return_info=False):
    """perturbations.py:1000-1061 (+ evolve_modes_batched :820-922): consecutive k-modes are grouped into batches of
    ``batch_size``; a batch starts at the smallest of its modes' start times (:786-805) and advances with one shared
    adaptive step.  Returns ``(y[num_k, nout, 20], kmodes)``.  (The reference reshapes ``ys[n_batches, nout, batch, n]``
    to ``(num_k, nout, n)`` without transposing, :907-909, which scrambles modes and output times whenever nout > 1;
    the modes are returned in order here.)"""
    if kmodes is None:
        kmodes = np.geomspace(kmin, kmax, num_k)
    kmodes = np.asarray(kmodes, dtype=np.float64)
    M = kmodes.shape[0]
    if M % batch_size != 0:
        raise ValueError("num_k must be divisible by batch_size")
    aexp_out = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    tau_out = param["tau_of_a_spline"].evaluate(aexp_out)
    tau_max = np.max(tau_out)
    d = Dims(lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax)
    nout = tau_out.shape[0]
    tau_start_modes = 0.99 * np.minimum(np.min(tau_out), determine_starting_time(param, kmodes))
    nb = M // batch_size
    yfull = np.zeros((M, nout, d.n))
    info = dict(nsteps=np.zeros(nb, dtype=np.int64), naccept=np.zeros(nb, dtype=np.int64), tau_start=np.zeros(nb), traces=[])
    for b in range(nb):
        sl = slice(b * batch_size, (b + 1) * batch_size)
        ts_b = float(np.min(tau_start_modes[sl]))
        y0 = adiabatic_ics(np.full(batch_size, ts_b), param, kmodes[sl], d)
        tr = []
        ys, st, ns, na = integrate_batch(ts_b, tau_max, y0, tau_out, param, kmodes[sl], d, rtol, atol, pcoeff, icoeff, dcoeff,
                                         factormax, factormin, max_steps, trace=tr)
        if st != 0:
            raise RuntimeError(f"oracle (batched): batch {b} failed with status {st}")
        yfull[sl] = ys
        info["nsteps"][b], info["naccept"][b], info["tau_start"][b] = ns, na, ts_b
        info["traces"].append(tr)
    y = convert_to_output(yfull, param, kmodes[:, None], d)
    if return_info:
        info["yfull"] = yfull
        return y, kmodes, info
    return y, kmodes
