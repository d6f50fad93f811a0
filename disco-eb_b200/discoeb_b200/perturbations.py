"""Drop-in host API of the Einstein-Boltzmann hot path, backed by the sm_100a CUDA library.

Mirrors the reference's keyword API on the ``param`` dict
(``/root/reference/src/discoeb/perturbations.py``):

* ``evolve_perturbations``          -> :926-997   returns ``(y, kmodes, param)``
* ``evolve_perturbations_batched``  -> :1000-1061 returns ``(y, kmodes)``
* ``get_power``                     -> :1101-1123   (and the other spectra entry points :1063-1224, see spectra.py)

Same argument names, defaults, return shapes, ``param`` side effects (:989-995) and failure
behaviour (diffrax raises when ``max_steps`` is exhausted; so does this).  The arithmetic runs in
``libdiscoeb_b200.so`` through the C-ABI of ``include/discoeb_b200.h``; there is no CPU path.

``evolve_perturbations_batched`` reproduces the reference's batched numerics: consecutive groups of
``batch_size`` modes start at the earliest of their start times and share ONE adaptive step size, controlled
by the RMS error over the whole batch (``Rodas5Batched``, ``ode_integrators_stiff.py:846-1010``;
``evolve_modes_batched``, ``perturbations.py:786-922``) -- the batched CUDA kernel keeps the warps of a batch in
lock-step through a thread-block cluster.  ``shared_step=False`` selects per-mode stepping instead (the
numerics of ``evolve_perturbations``, faster).  One reference quirk is consciously not reproduced: for more
than one output time the reference reshapes ``(n_batches, nout, batch, n)`` to ``(num_k, nout, n)`` without
transposing (:907-909), which scrambles modes and output times; the modes come back in order here.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from ._pack import pack_param, pack_params, pack_tangent

__all__ = ["evolve_perturbations", "evolve_perturbations_batched", "evolve_perturbations_multi", "evolve_perturbations_jvp",
           "get_power", "get_power_smoothed", "power_Kaiser", "power_multipoles", "get_xi_from_P"]


class MaxStepsReached(RuntimeError):
    """Same condition diffrax reports as ``RESULTS.max_steps_reached``."""


def _kgrid(kmin, kmax, num_k, dologk):
    # jnp.geomspace / jnp.linspace (perturbations.py:967-970)
    return np.geomspace(kmin, kmax, num_k) if dologk else np.linspace(kmin, kmax, num_k)


def _check_status(status, nsteps, max_steps, throw):
    if not throw:
        return
    bad = np.argwhere(status != 0)
    if bad.size:
        kinds = sorted(set(int(s) for s in status[status != 0]))
        msg = (f"{bad.shape[0]} of {status.size} modes did not complete (status codes {kinds}; 1 = the maximum "
               f"number of solver steps ({max_steps}) was reached, 2 = non-finite step, 3 = the mode was never integrated "
               f"(library fault)). Try increasing max_steps.")
        raise MaxStepsReached(msg)


def _solve(params, kmodes, aexp_out, *, lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax, rtol, atol, pcoeff, icoeff, dcoeff,
           factormax, factormin, max_steps, return_full, device, power_idx=-1, lib=None, k_per_cosmo=False, batch_size=0):
    lib = lib or _cabi.default_library()
    scalars, tables, nth, nnu = pack_params(params)
    aexp_out = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    if aexp_out.ndim != 1 or aexp_out.size < 1:
        raise ValueError("aexp_out must be a scalar or 1-d array")
    if np.any(np.diff(aexp_out) < 0):
        raise ValueError("aexp_out must be ascending (diffrax SaveAt(ts) requires increasing ts)")
    nk = kmodes.shape[-1]
    dims = _cabi.make_dims(ncosmo=len(params), nk=nk, nout=aexp_out.size, lmaxg=lmaxg, lmaxgp=lmaxgp, lmaxr=lmaxr,
                           lmaxnu=lmaxnu, nqmax=nqmax, nth=nth, nnu=nnu, max_steps=max_steps,
                           return_full=return_full, k_per_cosmo=k_per_cosmo, power_idx=power_idx, batch_size=batch_size)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=atol, pcoeff=pcoeff, icoeff=icoeff, dcoeff=dcoeff, factormax=factormax,
                           factormin=factormin)
    return lib.evolve_host(dims, ctrl, scalars, tables, kmodes, aexp_out, device=device, want_pk=power_idx >= 0)


def evolve_perturbations(*, param, aexp_out, kmin: float, kmax: float, num_k: int,
                         lmaxg: int = 11, lmaxgp: int = 11, lmaxr: int = 11, lmaxnu: int = 8,
                         nqmax: int = 3, rtol: float = 1e-4, atol: float = 1e-4,
                         pcoeff: float = 0.25, icoeff: float = 0.80, dcoeff: float = 0.0,
                         factormax: float = 20.0, factormin: float = 0.3, max_steps: int = 2048,
                         return_full: bool = False, dologk: bool = True, device: int = 0, throw: bool = True,
                         return_info: bool = False):
    """Evolve the linear perturbations of every k-mode in the synchronous gauge.

    Arguments, defaults and return value follow ``perturbations.py:926-997``:
    ``(y[num_k, nout, 20 | n], kmodes[num_k], param)`` with ``lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax,
    nout, tau_out`` written into ``param``.  Extra keyword-only knobs (not in the reference):
    ``device`` (CUDA ordinal), ``throw`` (False: return instead of raising on exhausted
    ``max_steps``) and ``return_info`` (append a dict with per-mode ``status``, ``nsteps``,
    ``naccept`` and the kernel time).
    """
    kmodes = _kgrid(kmin, kmax, num_k, dologk)
    out = _solve([param], kmodes, aexp_out, lmaxg=lmaxg, lmaxgp=lmaxgp, lmaxr=lmaxr, lmaxnu=lmaxnu, nqmax=nqmax,
                 rtol=rtol, atol=atol, pcoeff=pcoeff, icoeff=icoeff, dcoeff=dcoeff, factormax=factormax,
                 factormin=factormin, max_steps=max_steps, return_full=return_full, device=device)
    _check_status(out["status"], out["nsteps"], max_steps, throw)
    param["lmaxg"] = lmaxg
    param["lmaxgp"] = lmaxgp
    param["lmaxr"] = lmaxr
    param["lmaxnu"] = lmaxnu
    param["nqmax"] = nqmax
    param["nout"] = out["tau_out"].shape[1]
    param["tau_out"] = out["tau_out"][0]
    if return_info:
        info = dict(status=out["status"][0], nsteps=out["nsteps"][0], naccept=out["naccept"][0],
                    kernel_ms=out["kernel_ms"])
        return out["y"][0], kmodes, param, info
    return out["y"][0], kmodes, param


def evolve_perturbations_batched(*, param, aexp_out, kmin: float, kmax: float, num_k: int,
                                 lmaxg: int = 11, lmaxgp: int = 11, lmaxr: int = 11, lmaxnu: int = 8,
                                 nqmax: int = 3, rtol: float = 1e-4, atol: float = 1e-4,
                                 pcoeff: float = 0.25, icoeff: float = 0.80, dcoeff: float = 0.0,
                                 factormax: float = 20.0, factormin: float = 0.3, max_steps: int = 2048,
                                 batch_size: int = 16, device: int = 0, throw: bool = True, shared_step: bool = True):
    """``perturbations.py:1000-1061``: returns ``(y, kmodes)``; every batch of ``batch_size`` consecutive modes
    advances with one shared step size (see module docstring).  ``batch_size`` must divide ``num_k`` (:830) and be
    at most 64."""
    if num_k % batch_size != 0:
        raise ValueError("num_k must be divisible by batch_size (jnp.split at perturbations.py:830)")
    if shared_step and batch_size > 64:
        raise ValueError("batch_size > 64 is not supported by the shared-step kernel (use shared_step=False)")
    kmodes = np.geomspace(kmin, kmax, num_k)
    out = _solve([param], kmodes, aexp_out, lmaxg=lmaxg, lmaxgp=lmaxgp, lmaxr=lmaxr, lmaxnu=lmaxnu, nqmax=nqmax,
                 rtol=rtol, atol=atol, pcoeff=pcoeff, icoeff=icoeff, dcoeff=dcoeff, factormax=factormax,
                 factormin=factormin, max_steps=max_steps, return_full=False, device=device,
                 batch_size=batch_size if shared_step else 0)
    _check_status(out["status"], out["nsteps"], max_steps, throw)
    param["nout"] = out["tau_out"].shape[1]
    return out["y"][0], kmodes


def evolve_perturbations_multi(*, params, aexp_out, kmin: float, kmax: float, num_k: int,
                               lmaxg: int = 11, lmaxgp: int = 11, lmaxr: int = 11, lmaxnu: int = 8,
                               nqmax: int = 3, rtol: float = 1e-4, atol: float = 1e-4,
                               pcoeff: float = 0.25, icoeff: float = 0.80, dcoeff: float = 0.0,
                               factormax: float = 20.0, factormin: float = 0.3, max_steps: int = 2048,
                               return_full: bool = False, dologk: bool = True, device: int = 0, throw: bool = True,
                               power_idx: int = -1):
    """Batch of cosmologies in ONE launch (what ``jax.vmap(evolve_perturbations)`` over a stack of
    ``param`` dicts does in the reference's emulator/MCMC use).  Returns
    ``(y[ncosmo, num_k, nout, 20|n], kmodes, info)``; ``info['pk']`` holds ``get_power`` of field
    ``power_idx`` when requested (fused epilogue)."""
    kmodes = _kgrid(kmin, kmax, num_k, dologk)
    out = _solve(list(params), kmodes, aexp_out, lmaxg=lmaxg, lmaxgp=lmaxgp, lmaxr=lmaxr, lmaxnu=lmaxnu, nqmax=nqmax,
                 rtol=rtol, atol=atol, pcoeff=pcoeff, icoeff=icoeff, dcoeff=dcoeff, factormax=factormax,
                 factormin=factormin, max_steps=max_steps, return_full=return_full, device=device,
                 power_idx=power_idx)
    _check_status(out["status"], out["nsteps"], max_steps, throw)
    return out["y"], kmodes, out


def evolve_perturbations_jvp(*, param, dparam, aexp_out, kmin: float, kmax: float, num_k: int,
                             lmaxg: int = 11, lmaxgp: int = 11, lmaxr: int = 11, lmaxnu: int = 8,
                             nqmax: int = 3, rtol: float = 1e-4, atol: float = 1e-4,
                             pcoeff: float = 0.25, icoeff: float = 0.80, dcoeff: float = 0.0,
                             factormax: float = 20.0, factormin: float = 0.3, max_steps: int = 2048,
                             return_full: bool = False, dologk: bool = True, device: int = 0, throw: bool = True,
                             power_idx: int = -1, dkmin=0.0, dkmax=0.0):
    """Forward-mode derivative of ``evolve_perturbations`` -- what ``jax.jvp(evolve_perturbations, (param,),
    (dparam,))`` / ``jax.jacfwd`` returns in the reference (minimal notebook cell 14, Fisher notebook cell 7):
    the exact tangent of the discrete solve, computed by the tangent kernel in the same launch.

    ``dparam`` is one tangent of the ``param`` pytree (a dict with any subset of the scalar keys and of the spline
    keys, each spline carrying the tangents of ``x, y, S``), or a sequence of such dicts (the columns of a
    Jacobian).  Returns ``(y, dy, kmodes, info)``: ``dy`` has the shape of ``y`` (one dict) or a leading axis
    over directions; ``info`` holds ``tau_out, dtau_out, status, nsteps, naccept, kernel_ms`` and, with
    ``power_idx >= 0``, ``pk`` / ``dpk`` (``get_power`` and its tangent, A_s / n_s / k_p seeds included).
    ``dkmin`` / ``dkmax`` (a float, or one per direction) are the tangents of ``kmin`` / ``kmax`` for callers whose
    k grid moves with a parameter (k in units of h: ``nb_discoeb_rsd_eyes_plot.ipynb`` cell 5); the tangent of the
    ``geomspace`` / ``linspace`` grid follows.  ``param`` receives the same side effects as in ``evolve_perturbations``."""
    single = isinstance(dparam, dict)
    dlist = [dparam] if single else list(dparam)
    if not dlist:
        raise ValueError("dparam must hold at least one direction")
    lib = _cabi.default_library()
    kmodes = _kgrid(kmin, kmax, num_k, dologk)
    scalars, tables, nth, nnu = pack_params([param])
    seeds = [pack_tangent(param, d) for d in dlist]
    d_scalars = np.stack([s[0] for s in seeds])[:, None]
    d_tables = np.stack([s[1] for s in seeds])[:, None]
    aexp_out = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    if aexp_out.ndim != 1 or aexp_out.size < 1 or np.any(np.diff(aexp_out) < 0):
        raise ValueError("aexp_out must be a scalar or an ascending 1-d array")
    dims = _cabi.make_dims(ncosmo=1, nk=num_k, nout=aexp_out.size, lmaxg=lmaxg, lmaxgp=lmaxgp, lmaxr=lmaxr, lmaxnu=lmaxnu,
                           nqmax=nqmax, nth=nth, nnu=nnu, max_steps=max_steps, return_full=return_full, power_idx=power_idx,
                           ntan=len(dlist))
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=atol, pcoeff=pcoeff, icoeff=icoeff, dcoeff=dcoeff, factormax=factormax,
                           factormin=factormin)
    d_kmodes = None
    dkmin = np.broadcast_to(np.asarray(dkmin, dtype=np.float64), (len(dlist),))
    dkmax = np.broadcast_to(np.asarray(dkmax, dtype=np.float64), (len(dlist),))
    if np.any(dkmin != 0.0) or np.any(dkmax != 0.0):
        f = np.linspace(0.0, 1.0, num_k)
        if dologk:      # log k_i = (1 - f_i) log kmin + f_i log kmax
            d_kmodes = kmodes[None, :] * ((1.0 - f)[None, :] * (dkmin / kmin)[:, None] + f[None, :] * (dkmax / kmax)[:, None])
        else:
            d_kmodes = (1.0 - f)[None, :] * dkmin[:, None] + f[None, :] * dkmax[:, None]
    out = lib.evolve_tangent_host(dims, ctrl, scalars, tables, kmodes, aexp_out, d_scalars, d_tables, device=device,
                                  want_pk=power_idx >= 0, d_kmodes=d_kmodes)
    _check_status(out["status"], out["nsteps"], max_steps, throw)
    param["lmaxg"], param["lmaxgp"], param["lmaxr"], param["lmaxnu"], param["nqmax"] = lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax
    param["nout"] = out["tau_out"].shape[1]
    param["tau_out"] = out["tau_out"][0]
    dy = out["dy"][:, 0]
    info = dict(tau_out=out["tau_out"][0], dtau_out=out["dtau_out"][:, 0], status=out["status"][0], nsteps=out["nsteps"][0],
                naccept=out["naccept"][0], kernel_ms=out["kernel_ms"])
    if power_idx >= 0:
        info["pk"] = out["pk"][0]
        info["dpk"] = out["dpk"][0, 0] if single else out["dpk"][:, 0]
    return out["y"][0], (dy[0] if single else dy), kmodes, info


from .spectra import get_power, get_power_smoothed, power_Kaiser, power_multipoles, get_xi_from_P  # noqa: E402,F401
