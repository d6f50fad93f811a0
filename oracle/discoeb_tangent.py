"""CPU ORACLE for the forward-tangent row of the hot path --
TEST INFRASTRUCTURE.

What the reference computes when a caller wraps ``evolve_perturbations`` in ``jax.jvp`` /
``jax.jacfwd`` (``/root/reference/notebooks/nb_minimal_example.ipynb`` cell 14, the Fisher notebook
cell 7) is the exact derivative of the *discrete* solve: JAX pushes tangents through the start-time
bisection, the initial conditions, every Rosenbrock stage including the Jacobian and its LU
(``ode_integrators_stiff.py:772-779`` carry no ``stop_gradient``), the step end points
(``dt0 = min(t0/4, (t1-t0)/2)``, ``perturbations.py:756``; ``dt_next = dt * factor``), the
``SaveAt`` interpolation weights and the output conversion -- while every *decision* (accept/reject,
the PID factor, ``searchsorted`` intervals, bisection branches, clip-to-end) is taken on primal
values (diffrax marks the factor non-differentiable).

This module obtains that derivative without JAX by the complex-step method applied to the whole
restated algorithm of ``oracle/discoeb_oracle.py``: inputs become ``x + i h xdot`` with h = 1e-20,
all arithmetic is holomorphic, every comparison looks at real parts only, and the tangent is
``imag / h`` -- exact to round-off, no truncation error, no step-size tuning.  The one place the
primal oracle itself uses a complex step (the scale-factor column of the Jacobian) is replaced by
the hand-derived ``rhs_da`` there.

PARITY UNPINNED (same reason as the primal oracle: JAX/diffrax are not installable here); what
pins it: (i) it *is* the primal oracle's code path, (ii) central finite differences of the primal
oracle along a frozen step sequence agree (tests/test_oracle_tangent.py).

Only ``tests/``, ``tools/`` fixture generators and ``__graft_entry__.smoke()`` may import this file.
"""
from __future__ import annotations

import numpy as np

from . import discoeb_oracle as O
from .background import Spline

H_CS = 1e-20

SCALAR_KEYS = ("Omegam", "Omegab", "OmegaDE", "Omegak", "grhom", "grhog", "grhor", "Neff", "Nmnu", "amnu",
               "w_DE_0", "w_DE_a", "cs2_DE", "YHe", "H0", "taumin", "A_s", "n_s", "k_p")
SPLINE_KEYS = ("cs2a_of_loga_spline", "xe_of_loga_spline", "logrhonu_of_loga_spline", "logpnu_of_loga_spline",
               "a_of_tau_spline", "xe_of_tau_spline", "tau_of_a_spline")


def complexify(param, dparam, h=H_CS):
    """param + i h dparam.  ``dparam`` maps scalar keys to floats and spline keys to objects with
    ``x, y, S`` tangent arrays (missing keys = zero tangent)."""
    pc = {}
    for key in SCALAR_KEYS:
        if key in param:
            pc[key] = complex(float(param[key]), h * float(dparam.get(key, 0.0)))
    for key in SPLINE_KEYS:
        sp = param[key]
        c = Spline.__new__(Spline)
        if key in dparam:
            ds = dparam[key]
            c.x = sp.x + 1j * h * np.asarray(ds.x)
            c.y = sp.y + 1j * h * np.asarray(ds.y)
            c.S = sp.S + 1j * h * np.asarray(ds.S)
        else:
            c.x, c.y, c.S = sp.x.astype(np.complex128), sp.y.astype(np.complex128), sp.S.astype(np.complex128)
        pc[key] = c
    return pc


def _where_le(a, b):
    """min(a, b) decided on real parts."""
    return np.where(np.real(a) <= np.real(b), a, b)


def integrate_modes_cs(t0, t1, y0, ts, p, k, d, rtol, atol, pcoeff=0.25, icoeff=0.8, dcoeff=0.0, factormax=20.0,
                       factormin=0.3, max_steps=2048, safety=0.9, order=5.0, trace=None, replay=None):
    """``discoeb_oracle.integrate_modes`` with complex state/times; decisions on real parts.

    ``replay`` = (keep[M, S], fac[M, S], n[M]) freezes the accept/reject decisions and step-size factors to those
    of an earlier run (used by the finite-difference cross-check)."""
    M, n = y0.shape
    nout = len(ts)
    ys = np.full((M, nout, n), np.nan + 0j)
    tprev = np.array(t0, dtype=np.complex128)
    t1 = complex(t1)
    ts = np.asarray(ts, dtype=np.complex128)
    dt0 = _where_le(tprev / 4, 0.5 * (t1 - tprev))
    tnext = tprev + dt0
    tnext = np.where(np.real(tnext) > np.real(t1) - 1e-10, t1, tnext)
    y = np.array(y0, dtype=np.complex128)
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
        lim = max_steps if replay is None else replay[2]
        act = np.nonzero((np.real(tprev) < np.real(t1)) & (status == 0) & (nsteps < lim))[0]
        if act.size == 0:
            break
        tp, tn, ya, ka = tprev[act], tnext[act], y[act], k[act]
        with np.errstate(all="ignore"):
            y1, err = O.rodas5_step(tp, tn, ya, p, ka, d)
            errr = np.where(np.isnan(np.real(err)), np.inf, np.real(err))
            E = O.scaled_error_norm(np.real(ya), np.real(y1), errr, np.real(ka), rtol, atol)
            if replay is None:
                keep = E < 1
                inv = 1.0 / E
                f1 = inv ** c1 if c1 != 0 else 1.0
                f2 = inv_prev[act] ** c2 if c2 != 0 else 1.0
                f3 = inv_pprev[act] ** c3 if c3 != 0 else 1.0
                fac = np.clip(safety * f1 * f2 * f3, np.where(keep, 1.0, factormin), factormax)
                inv = np.where((inv == 0) | np.isinf(inv), 1.0, inv)
            else:
                keep = replay[0][act, nsteps[act]] != 0
                fac = replay[1][act, nsteps[act]]
                inv = np.ones(act.size)
            dtn = (tn - tp) * fac
        if trace is not None:
            trace.append((act.copy(), tp.copy(), tn.copy(), E.copy(), keep.copy(), np.array(fac, dtype=np.float64).copy()))
        nsteps[act] += 1
        nacc[act] += keep
        for m in np.nonzero(keep)[0]:
            g = act[m]
            while save_idx[g] < nout and np.real(ts[save_idx[g]]) <= np.real(tn[m]):
                tt = ts[save_idx[g]]
                coeff = 0.0 if tn[m] == tp[m] else (tt - tp[m]) / (tn[m] - tp[m])
                ys[g, save_idx[g]] = ya[m] + coeff * (y1[m] - ya[m])
                save_idx[g] += 1
        y[act] = np.where(keep[:, None], y1, ya)
        tpn = np.where(keep, tn, tp)
        tpn = np.where(np.real(tpn) <= np.real(t1), tpn, t1)
        tnn = tpn + dtn
        clip = np.real(tnn) > np.real(t1) - 1e-10
        tnn = np.where(clip, np.where(keep, t1, tpn + 0.5 * (t1 - tpn)), tnn)
        tprev[act], tnext[act] = tpn, tnn
        inv_pprev[act] = np.where(keep, inv_prev[act], inv_pprev[act])
        inv_prev[act] = np.where(keep, inv, inv_prev[act])
        bad = ~np.isfinite(np.real(tnn)) | (~np.isfinite(np.real(y[act]))).any(-1)
        status[act[bad]] = 2
    status[(status == 0) & (np.real(tprev) < np.real(t1))] = 1
    return ys, status, nsteps, nacc


def evolve_perturbations_jvp(*, param, dparam, aexp_out, kmodes, dkmodes=None, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3,
                             rtol=1e-4, atol=1e-4, pcoeff=0.25, icoeff=0.80, dcoeff=0.0, factormax=20.0, factormin=0.3,
                             max_steps=2048, h=H_CS, replay=None):
    """Primal and tangent of ``evolve_perturbations`` (perturbations.py:926-997) along ``dparam``.

    Returns a dict: ``y, dy`` [M, nout, 20]; ``yfull, dyfull`` [M, nout, n]; ``pk4/dpk4`` and ``pk6/dpk6``
    (get_power of fields 4 and 6 with its own A_s/n_s/k_p tangents); ``tau_out, dtau_out``; ``tau_start,
    dtau_start``; ``y0, dy0``; the step trace ``rp_tnext, rp_dtnext, rp_keep, rp_fac`` [M, S]; ``nsteps``,
    ``naccept``."""
    pc = complexify(param, dparam, h)
    kreal = np.asarray(kmodes, dtype=np.float64)
    # tangent of the wavenumbers themselves (callers that scale kmin/kmax by h: nb_discoeb_rsd_eyes_plot.ipynb cell 5)
    kmodes = kreal if dkmodes is None else kreal + 1j * h * np.asarray(dkmodes, dtype=np.float64)
    aexp_out = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    d = O.Dims(lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax)
    tau_out = pc["tau_of_a_spline"].evaluate(aexp_out)
    tau_max = tau_out[np.argmax(np.real(tau_out))]
    tau_min = tau_out[np.argmin(np.real(tau_out))]
    ts0 = O.determine_starting_time(pc, kmodes)
    tau_start = 0.99 * np.where(np.real(tau_min) <= np.real(ts0), tau_min, ts0)
    y0 = O.adiabatic_ics(tau_start, pc, kmodes, d)
    tr = []
    ys, st, ns, na = integrate_modes_cs(tau_start, tau_max, y0, tau_out, pc, kmodes, d, rtol, atol, pcoeff, icoeff, dcoeff,
                                        factormax, factormin, max_steps, trace=tr, replay=replay)
    if np.any(st != 0):
        raise RuntimeError(f"tangent oracle: {np.count_nonzero(st)} modes failed (status {np.unique(st)})")
    M = len(kmodes)
    S = int(ns.max())
    rp_t = np.zeros((M, S), dtype=np.complex128)
    rp_k = np.zeros((M, S), dtype=np.int32)
    rp_f = np.zeros((M, S))
    cnt = np.zeros(M, dtype=np.int64)
    for act, tp, tn, E, keep, fac in tr:
        rp_t[act, cnt[act]] = tn
        rp_k[act, cnt[act]] = keep
        rp_f[act, cnt[act]] = fac
        cnt[act] += 1
    y20 = O.convert_to_output(ys, pc, kmodes[:, None], d)
    out = dict(kmodes=kreal, aexp_out=aexp_out, nsteps=ns.astype(np.int32), naccept=na.astype(np.int32),
               rp_tnext=rp_t.real.copy(), rp_dtnext=rp_t.imag / h, rp_keep=rp_k, rp_fac=rp_f)
    for name, z in (("y", y20), ("yfull", ys), ("tau_out", tau_out), ("tau_start", tau_start), ("y0", y0)):
        out[name] = np.real(z).copy()
        out["d" + name] = np.imag(z) / h
    for idx in (4, 6):
        pk = O.get_power(k=kmodes[:, None], y=y20, idx=idx, param=pc)
        out[f"pk{idx}"] = np.real(pk).copy()
        out[f"dpk{idx}"] = np.imag(pk) / h
    return out
