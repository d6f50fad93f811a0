"""diffrax stand-in.  Test infrastructure.

The classes the reference subclasses or instantiates are plain Python here; `diffeqsolve` restates diffrax's adaptive
loop for the one configuration the reference uses it in: an ODETerm, an AbstractAdaptiveSolver with a
LocalLinearInterpolation dense output, `PIDController`, `SaveAt(ts=...)` (or t1=True), no events, no jumps.
Semantics follow diffrax 0.6/0.7 (`_integrate.py: loop/body_fun/_clip_to_end`, `_step_size_controller/adaptive.py:
PIDController.init/adapt_step_size`, `_local_interpolation.py`, `_misc.py: linear_rescale`):

  * first step  tnext = min(t0 + dt0, t1); controller state (prev_inv_err, prev_prev_inv_err) = (1, 1);
  * per attempted step: (y1, y_error) = solver.step(tprev, tnext); NaNs in y_error -> inf;
    scaled = y_error / (atol + max(|y0|, |y1 or y0 if y1 has a NaN|) rtol);  E = norm(scaled);  keep = E < 1;
    factor = clip(safety * (1/E)^c1 * prev^c2 * prevprev^c3, keep ? 1 : factormin, factormax) with
    c1 = (i+p+d)/order, c2 = -(p+2d)/order, c3 = d/order, order = solver.error_order(terms) (= solver.order for ODEs
    unless the solver overrides it);  dt = (tnext - tprev) * factor;  1/E of 0 or inf is stored as 1 and the history
    shifts on accepted steps only;
  * next interval: from tnext if kept else from tprev; clipped to the end: if tnext' > t1 - 1e-10 (float64) then
    tnext' = t1 if kept else tprev + (t1 - tprev)/2;
  * SaveAt(ts): on a kept step every ts[j] <= tnext (in order) is sampled by linear interpolation between the step's
    end states, coefficient (ts[j]-tprev)/(tnext-tprev) (0 when the interval is empty);
  * steps (kept or not) count against max_steps; running out raises unless throw=False.
"""
import numpy as _np
import jax.numpy as jnp
from jax.numpy import _wrap


class RESULTS:
    successful = 0
    max_steps_reached = 1
    dt_min_reached = 2
    nan_time = 3


class AbstractTerm:
    pass


class ODETerm(AbstractTerm):
    def __init__(self, vector_field):
        self.vector_field = vector_field

    def vf(self, t, y, args):
        return self.vector_field(t, y, args)

    def contr(self, t0, t1, **kw):
        return t1 - t0

    def prod(self, vf, control):
        return vf * control

    def vf_prod(self, t, y, args, control):
        return self.vf(t, y, args) * control


class AbstractSolver:
    def error_order(self, terms):
        return self.order(terms)

    def init(self, terms, t0, t1, y0, args):
        return None


class AbstractAdaptiveSolver(AbstractSolver):
    pass


class AbstractImplicitSolver(AbstractSolver):
    pass


class LocalLinearInterpolation:
    def __init__(self, *, t0, t1, y0, y1, **kw):
        self.t0, self.t1, self.y0, self.y1 = t0, t1, y0, y1

    def evaluate(self, t0, t1=None, left=True):
        cond = self.t0 == self.t1
        div = 1.0 if cond else self.t1 - self.t0
        coeff = (t0 - self.t0) / div
        return self.y0 + coeff * (self.y1 - self.y0)


class SaveAt:
    def __init__(self, *, t0=False, t1=False, ts=None, steps=False, dense=False, fn=None, subs=None, solver_state=False,
                 controller_state=False, made_jump=False):
        self.t0, self.t1, self.ts, self.steps, self.dense = t0, t1, ts, steps, dense


class PIDController:
    def __init__(self, rtol, atol, pcoeff=0.0, icoeff=1.0, dcoeff=0.0, dtmin=None, dtmax=None, force_dtmin=True, step_ts=None,
                 jump_ts=None, factormin=0.2, factormax=10.0, norm=None, safety=0.9, error_order=None):
        self.rtol, self.atol, self.pcoeff, self.icoeff, self.dcoeff = rtol, atol, pcoeff, icoeff, dcoeff
        self.dtmin, self.dtmax, self.factormin, self.factormax = dtmin, dtmax, factormin, factormax
        self.norm = norm if norm is not None else (lambda x: _np.sqrt(_np.mean(_np.asarray(x) ** 2)))
        self.safety, self.error_order = safety, error_order
        assert step_ts is None and jump_ts is None and dtmin is None, "not needed by the reference's calls"


class _Adjoint:
    def __init__(self, *a, **k):
        pass


DirectAdjoint = RecursiveCheckpointAdjoint = BacksolveAdjoint = ImplicitAdjoint = ForwardMode = _Adjoint


class Kvaerno5(AbstractAdaptiveSolver):
    def __init__(self, *a, **k):
        raise NotImplementedError("refshim: Kvaerno5 is not on the reference's default path")


class Tsit5(AbstractAdaptiveSolver):
    def __init__(self, *a, **k):
        raise NotImplementedError


class Solution:
    def __init__(self, ts, ys, stats, result, trace=None):
        self.ts, self.ys, self.stats, self.result, self.trace = ts, ys, stats, result, trace


TRACE = None          # set to a list to record (tprev, tnext, E, keep) of every attempted step


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, saveat=None, stepsize_controller=None, adjoint=None,
                max_steps=4096, throw=True, **kw):
    ctl = stepsize_controller
    saveat = saveat if saveat is not None else SaveAt(t1=True)
    t0 = float(_np.real(t0)); t1 = float(_np.real(t1))
    order = ctl.error_order if ctl.error_order is not None else solver.error_order(terms)
    c1 = (ctl.icoeff + ctl.pcoeff + ctl.dcoeff) / order
    c2 = -(ctl.pcoeff + 2 * ctl.dcoeff) / order
    c3 = ctl.dcoeff / order
    tprev, tnext = t0, min(t0 + float(_np.real(dt0)), t1)
    if ctl.dtmax is not None:
        tnext = min(tnext, tprev + ctl.dtmax)
    y = _wrap(_np.array(y0, dtype=float))
    solver_state = solver.init(terms, t0, t1, y0, args)
    prev_inv, pprev_inv = 1.0, 1.0
    ts = None if saveat.ts is None else _np.atleast_1d(_np.asarray(saveat.ts, dtype=float))
    ys = [] if ts is None else [None] * len(ts)
    save_idx, nsteps, naccept, result = 0, 0, 0, RESULTS.successful
    trace = []
    while tprev < t1 and nsteps < max_steps:
        y1, y_error, dense_info, solver_state, _ = solver.step(terms, tprev, tnext, y, args, solver_state, False)
        y1 = _np.asarray(y1); y_error = _np.asarray(y_error)
        y_error = _np.where(_np.isnan(y_error), _np.inf, y_error)
        ycand = _np.asarray(y) if _np.isnan(y1).any() else y1
        scale = ctl.atol + _np.maximum(_np.abs(_np.asarray(y)), _np.abs(ycand)) * ctl.rtol
        E = float(ctl.norm(_wrap(y_error / scale)))
        keep = E < 1
        with _np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            inv = 1.0 / E if E != 0 else _np.inf
            f1 = 1.0 if c1 == 0 else inv ** c1
            f2 = 1.0 if c2 == 0 else prev_inv ** c2
            f3 = 1.0 if c3 == 0 else pprev_inv ** c3
            factor = float(_np.clip(ctl.safety * f1 * f2 * f3, 1.0 if keep else ctl.factormin, ctl.factormax))
        dt = (tnext - tprev) * factor
        if ctl.dtmax is not None:
            dt = min(dt, ctl.dtmax)
        if inv == 0 or _np.isinf(inv):
            inv = 1.0
        trace.append((tprev, tnext, E, bool(keep)))
        nsteps += 1
        if keep:
            naccept += 1
            interp = solver.interpolation_cls(t0=tprev, t1=tnext, **dense_info)
            while ts is not None and save_idx < len(ts) and ts[save_idx] <= tnext:
                ys[save_idx] = _np.asarray(interp.evaluate(ts[save_idx]))
                save_idx += 1
            pprev_inv, prev_inv = prev_inv, inv
            y = _wrap(_np.array(y1))
            tprev = min(tnext, t1)
        tn = tprev + dt
        if tn > t1 - 1e-10:
            tn = t1 if keep else tprev + 0.5 * (t1 - tprev)
        tnext = tn
        if not _np.isfinite(tnext):
            result = RESULTS.nan_time
            break
    if result == RESULTS.successful and tprev < t1:
        result = RESULTS.max_steps_reached
    if result != RESULTS.successful and throw:
        raise RuntimeError(f"refshim.diffeqsolve: result {result} after {nsteps} steps (max_steps={max_steps})")
    if ts is None:
        out_t, out_y = _np.array([tprev]), _wrap(_np.asarray(y)[None])
    else:
        n = _np.asarray(y).shape[0]
        out_t = ts
        out_y = _wrap(_np.stack([v if v is not None else _np.full(n, _np.inf) for v in ys]))
    if TRACE is not None:
        TRACE.append(trace)
    return Solution(_wrap(out_t), out_y, dict(num_steps=nsteps, num_accepted_steps=naccept, num_rejected_steps=nsteps - naccept),
                    result, trace)
