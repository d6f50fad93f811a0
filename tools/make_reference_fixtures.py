"""Runs the UNMODIFIED reference sources (/root/reference/src/discoeb) in this container and commits what they return.

JAX/diffrax are not installable here; ``tools/refshim`` provides NumPy stand-ins (README there) under which the
reference's own Python executes statement by statement: ``model_synchronous``, ``Rodas5Transformed.step`` (with
``jax.jacfwd`` as an exact complex-step Jacobian and LAPACK LU), ``determine_starting_time``,
``adiabatic_ics_one_mode``, ``evolve_one_mode`` (through the restated diffrax loop), ``convert_to_output_variables``,
``get_power`` and the ``spline_interpolation`` constructor.  Output: ``tests/golden/ref_<case>.npz`` in the layout of
``oracle_<case>.npz`` (tools/make_golden.py) plus single-function vectors, so that

* the CPU suite pins the oracle (oracle/discoeb_oracle.py) against the reference itself (tests/test_reference_pin.py),
* the GPU suite replays the kernel along the REFERENCE's step sequence and compares with the REFERENCE's outputs.

Only this script reads /root/reference; the fixtures travel, the reference does not.

    python tools/make_reference_fixtures.py [case ...]      (8 worker processes; minutes per n=265 case)
"""
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))      # shadows jaxtyping too
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore", category=SyntaxWarning)

import jax  # noqa: E402  (refshim)
import jax.numpy as jnp  # noqa: E402
import diffrax  # noqa: E402  (refshim)
from discoeb import perturbations as RP  # noqa: E402  (the reference)
from discoeb.ode_integrators_stiff import Rodas5Transformed  # noqa: E402
from discoeb.spline_interpolation import spline_interpolation as RSpline  # noqa: E402
import helpers  # noqa: E402
from discoeb_b200._pack import SCALAR_KEYS, SPLINE_KEYS  # noqa: E402
from make_golden import CASES  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# extra reference-run cases next to the twins of tools/make_golden.py's CASES
EXTRA = {
    # BASELINE config 2 (n=265, z=0): every 32nd mode of the 512-mode bench grid, the ~600-step k=10 mode included
    "config2_grid_n265": ("fiducial", (31, 31, 31, 31, 5), np.geomspace(1e-4, 10.0, 512)[31::32], [1.0], 1e-4),
    # BASELINE config 3 (w0wa + massive nu, n=265): every 512th mode of the 4096-mode grid
    "config3_grid_n265": ("w0wa", (31, 31, 31, 31, 5), np.geomspace(1e-4, 10.0, 4096)[255::512], [1.0], 1e-4),
}


def reference_param(name):
    """The reference's ``param`` dict for a committed table set: scalars + splines built by the REFERENCE's constructor
    from the knots (its second derivatives are compared with the stored ones below)."""
    po = helpers.load_tables(name).param()
    pr = {k: po[k] for k in SCALAR_KEYS}
    worst = 0.0
    for key in SPLINE_KEYS:
        sp = RSpline(jnp.array(po[key].x), jnp.array(po[key].y))
        worst = max(worst, float(np.abs(np.asarray(sp._S_full_) - po[key].S).max() / np.abs(po[key].S).max()))
        pr[key] = sp
    return pr, worst


_G = {}


def _one_mode(i):
    pr, dims, ks, tau_out, rtol, max_steps = _G["args"]
    lg, lp, lr, ln, nq = dims
    k = float(ks[i])
    diffrax.TRACE = []
    t = time.time()
    yfull = RP.evolve_one_mode(tau_max=float(np.max(tau_out)), tau_out=jnp.array(tau_out), param=pr, kmode=k, lmaxg=lg, lmaxgp=lp,
                               lmaxr=lr, lmaxnu=ln, nqmax=nq, rtol=rtol, atol=rtol, pcoeff=0.25, icoeff=0.80, dcoeff=0.0,
                               factormax=20.0, factormin=0.3, max_steps=max_steps, return_full=True)
    trace = diffrax.TRACE[-1]
    y20 = jax.vmap(lambda y: RP.convert_to_output_variables(y=y, param=pr, kmode=k, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln,
                                                            nqmax=nq))(yfull)
    tau_start = 0.99 * min(float(np.min(tau_out)), float(RP.determine_starting_time(param=pr, k=k)))
    nvar = 7 + (lg + 1) + (lp + 1) + (lr + 1) + nq * (ln + 1) + 2
    y0 = RP.adiabatic_ics_one_mode(tau=tau_start, param=pr, kmode=k, nvar=nvar, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq)
    # single-function vectors: RHS at a scrambled state, one Rodas5 step from the initial state with dt0
    rng = np.random.default_rng(1000 + i)
    ys = np.asarray(y0) * (1.0 + 0.3 * rng.normal(size=nvar)) + 1e-3 * np.abs(np.asarray(y0)).max() * rng.normal(size=nvar)
    ys[0] = float(y0[0])
    f_ref = RP.model_synchronous(tau=tau_start * 1.7, y=jnp.array(ys), param=pr, kmode=k, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln,
                                 nqmax=nq)
    term = diffrax.ODETerm(lambda tau, y, params: RP.model_synchronous(tau=tau, y=y, param=params[0], kmode=params[1], lmaxg=lg,
                                                                       lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq))
    t1s = tau_start + min(tau_start / 4, 0.5 * (float(np.max(tau_out)) - tau_start))
    y1, err, _, _, _ = Rodas5Transformed().step(term, tau_start, t1s, jnp.array(ys), (pr, k), None, False)
    print(f"  mode {i} k={k:.4g}: {len(trace)} steps, {time.time() - t:.1f}s", flush=True)
    return dict(yfull=np.asarray(yfull), y=np.asarray(y20), tau_start=tau_start, y0=np.asarray(y0), trace=trace,
                rhs_state=ys, rhs_tau=tau_start * 1.7, rhs_f=np.asarray(f_ref), step_t0=tau_start, step_t1=t1s, step_y1=np.asarray(y1),
                step_err=np.asarray(err))


def run_case(name, cos, dims, ks, aout, rtol, max_steps=4096, workers=8):
    pr, sdiff = reference_param(cos)
    ks = np.asarray(ks, dtype=np.float64)
    aout = np.asarray(aout, dtype=np.float64)
    tau_out = np.asarray(jax.vmap(lambda a: pr["tau_of_a_spline"].evaluate(a))(jnp.array(aout)))
    _G["args"] = (pr, dims, ks, tau_out, rtol, max_steps)
    order = np.argsort(-ks)                       # longest modes first
    with mp.get_context("fork").Pool(min(workers, len(ks))) as pool:
        res = pool.map(_one_mode, [int(i) for i in order], chunksize=1)
    res = [res[int(np.nonzero(order == i)[0][0])] for i in range(len(ks))]
    M = len(ks)
    ns = np.array([len(r["trace"]) for r in res], dtype=np.int32)
    stride = int(ns.max())
    rp_t = np.zeros((M, stride)); rp_k = np.zeros((M, stride), dtype=np.int32); rp_E = np.zeros((M, stride))
    for m, r in enumerate(res):
        for s_, (tp, tn, E, keep) in enumerate(r["trace"]):
            rp_t[m, s_], rp_k[m, s_], rp_E[m, s_] = tn, keep, E
    pk4 = np.asarray(RP.get_power(k=jnp.array(ks)[:, None], y=jnp.array(np.stack([r["y"] for r in res])), idx=4, param=pr))
    stack = lambda key: np.stack([r[key] for r in res])
    return dict(cosmology=cos, kmodes=ks, aexp_out=aout, tau_out=tau_out, tau_start=stack("tau_start"), y0=stack("y0"),
                yfull=stack("yfull"), y=stack("y"), pk4=pk4, nsteps=ns, naccept=rp_k.sum(1).astype(np.int32), rp_tnext=rp_t, rp_keep=rp_k,
                rp_E=rp_E, dims=np.array(dims), rtol=rtol, rhs_state=stack("rhs_state"), rhs_tau=stack("rhs_tau"), rhs_f=stack("rhs_f"),
                step_t0=stack("step_t0"), step_t1=stack("step_t1"), step_y1=stack("step_y1"), step_err=stack("step_err"),
                spline_S_maxdiff=sdiff, produced_by="reference sources under tools/refshim (make_reference_fixtures.py)")


def main():
    only = sys.argv[1:]
    allc = dict(CASES); allc.update(EXTRA)
    for name, (cos, dims, ks, aout, rtol) in allc.items():
        if only and name not in only:
            continue
        t = time.time()
        out = run_case(name, cos, dims, ks, aout, rtol)
        np.savez_compressed(os.path.join(GOLD, f"ref_{name}.npz"), **out)
        print(name, "steps", out["nsteps"].tolist(), "%.1fs" % (time.time() - t), flush=True)


if __name__ == "__main__":
    main()
