"""Pins the forward-tangent oracle (oracle/discoeb_tangent.py, complex step through the restated algorithm) against the
REFERENCE's own functions, executed from /root/reference/src through tools/refshim: central differences with one
Richardson step of

    reference ICs -> reference Rodas5Transformed.step along the recorded step sequence -> linear output interpolation
    -> reference convert_to_output_variables -> reference get_power

under param(eps) = param + eps * seed, with the step times, start time and output times moved along the oracle's own
tangents (t_i + eps * dt_i: the frozen step sequence of the replay tests).  Writes tests/golden/reference_tangent_<case>.npz:
the differenced tangents, their Richardson error estimates, and the reference's differenced start/output times.

TEST INFRASTRUCTURE: runs in the builder container only (needs /root/reference); the tests read the committed npz.
"""
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))
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

GOLD = os.path.join(ROOT, "tests", "golden")
# step of the differences relative to the size of the parameter a direction stands for
SCALE = {"Omegam": 0.3, "Omegab": 0.05, "H0": 70.0, "w_DE_0": 1.0, "w_DE_a": 1.0}
CASES = {
    # fixture: (modes, directions) differenced; directions that only enter the P(k) epilogue (n_s, A_s) are checked there
    "default_n72": ((0, 1, 2), ("Omegam", "H0", "w_DE_0")),
    "w0wa_n43": ((0, 1), ("w_DE_a", "Omegab")),
    "fisher_n265": ((0,), ("Omegam",)),
}
REL = 2e-4


def reference_param(scal, tab, nth, nnu):
    """The reference's param dict from packed inputs: its own spline objects, holding exactly the packed x, y, S."""
    po = helpers.Tables(scal, tab, nth, nnu).param()
    pr = {k: po[k] for k in SCALAR_KEYS if k in po}
    for key in SPLINE_KEYS:
        x, y, S = (np.asarray(v, dtype=np.float64) for v in (po[key].x, po[key].y, po[key].S))
        sp = RSpline(jnp.array(x), jnp.array(y))
        sp._x_, sp._y_, sp._S_full_ = jnp.array(x), jnp.array(y), jnp.array(S)
        pr[key] = sp
    return pr


_G = {}


def _replay(job):
    m, d, eps = job
    z, dims = _G["z"], _G["dims"]
    lg, lp, lr, ln, nq = dims
    pr = reference_param(z["scalars"] + eps * z["d_scalars"][d], z["tables"] + eps * z["d_tables"][d], z["nth"], z["nnu"])
    k = float(z["kmodes"][m]) + (eps * float(z["d_kmodes"][d, m]) if "d_kmodes" in z else 0.0)
    nvar = 7 + (lg + 1) + (lp + 1) + (lr + 1) + nq * (ln + 1) + 2
    t0 = float(z["tau_start"][m] + eps * z["dtau_start"][d, m])
    tout = z["tau_out"] + eps * z["dtau_out"][d]
    y = RP.adiabatic_ics_one_mode(tau=t0, param=pr, kmode=k, nvar=nvar, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq)
    y0 = np.asarray(y).copy()
    term = diffrax.ODETerm(lambda tau, yy, params: RP.model_synchronous(tau=tau, y=yy, param=params[0], kmode=params[1], lmaxg=lg,
                                                                        lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq))
    solver = Rodas5Transformed()
    t = t0
    ys = np.zeros((len(tout), nvar))
    si = 0
    for s in range(int(z["nsteps"][m])):
        tn = float(z["rp_tnext"][m, s] + eps * z["rp_dtnext"][d, m, s])
        y1, _, _, _, _ = solver.step(term, t, tn, y, (pr, k), None, False)
        if z["rp_keep"][m, s]:
            while si < len(tout) and tout[si] <= tn:          # SaveAt(ts): linear interpolation inside the accepted step
                c = 0.0 if tn == t else (tout[si] - t) / (tn - t)
                ys[si] = np.asarray(y) + c * (np.asarray(y1) - np.asarray(y))
                si += 1
            y, t = y1, tn
    assert si == len(tout), (m, d, eps, si)
    y20 = np.stack([np.asarray(RP.convert_to_output_variables(y=jnp.array(v), param=pr, kmode=k, lmaxg=lg, lmaxgp=lp, lmaxr=lr,
                                                              lmaxnu=ln, nqmax=nq)) for v in ys])
    pk4 = np.asarray(RP.get_power(k=jnp.array([k])[:, None], y=jnp.array(y20[None]), idx=4, param=pr))[0]
    # the reference's own start and output times of the shifted cosmology (differenced below against the oracle's tangents)
    ts_ref = 0.99 * min(float(np.min(tout)), float(RP.determine_starting_time(param=pr, k=k)))
    to_ref = np.asarray(jax.vmap(lambda a: pr["tau_of_a_spline"].evaluate(a))(jnp.array(z["aexp_out"])))
    return dict(yfull=ys, y=y20, pk4=pk4, y0=y0, tau_start=ts_ref, tau_out=to_ref)


def richardson(res, eps):
    """res[(sign, half)] -> (derivative, error estimate) per key"""
    out = {}
    for key in res[(1, 0)]:
        d1 = (np.asarray(res[(1, 0)][key]) - np.asarray(res[(-1, 0)][key])) / (2 * eps)
        d2 = (np.asarray(res[(1, 1)][key]) - np.asarray(res[(-1, 1)][key])) / eps
        out[key] = ((4 * d2 - d1) / 3, np.abs(d2 - d1) / 3)
    return out


def main():
    only = sys.argv[1:]
    workers = int(os.environ.get("WORKERS", "8"))
    for name, (modes, dirs) in CASES.items():
        if only and name not in only:
            continue
        t = time.time()
        z = dict(np.load(os.path.join(GOLD, f"tangent_{name}.npz"), allow_pickle=False))
        alld = [str(s) for s in z["directions"]]
        _G["z"], _G["dims"] = z, tuple(int(v) for v in z["dims"])
        jobs, meta = [], []
        for m in modes:
            for key in dirs:
                d = alld.index(key)
                eps = REL * SCALE[key]
                for sign in (1, -1):
                    for half in (0, 1):
                        jobs.append((m, d, sign * eps / (2 if half else 1)))
                        meta.append((m, d, sign, half))
        jobs_sorted = sorted(range(len(jobs)), key=lambda i: -int(z["nsteps"][jobs[i][0]]))
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            rs = pool.map(_replay, [jobs[i] for i in jobs_sorted], chunksize=1)
        res = {}
        for i, r in zip(jobs_sorted, rs):
            m, d, sign, half = meta[i]
            res.setdefault((m, d), {})[(sign, half)] = r
        save = dict(modes=np.array(modes), directions=np.array(dirs), dir_index=np.array([alld.index(k) for k in dirs]),
                    eps=np.array([REL * SCALE[k] for k in dirs]),
                    produced_by="central differences + Richardson of the reference's functions under tools/refshim (make_reference_tangent.py)")
        keys = ("yfull", "y", "pk4", "y0", "tau_start", "tau_out")
        acc = {k: [] for k in keys}; err = {k: [] for k in keys}
        for m in modes:
            rowa = {k: [] for k in keys}; rowe = {k: [] for k in keys}
            for key in dirs:
                d = alld.index(key)
                r = richardson(res[(m, d)], REL * SCALE[key])
                for k in keys:
                    rowa[k].append(r[k][0]); rowe[k].append(r[k][1])
            for k in keys:
                acc[k].append(np.stack(rowa[k])); err[k].append(np.stack(rowe[k]))
        for k in keys:          # [mode, direction, ...]
            save["d" + k] = np.stack(acc[k]); save["d" + k + "_err"] = np.stack(err[k])
        np.savez_compressed(os.path.join(GOLD, f"reference_tangent_{name}.npz"), **save)
        # report against the oracle's tangents (the tests apply the per-field scales of tests/parity_checks.py)
        for im, m in enumerate(modes):
            for idd, key in enumerate(dirs):
                d = alld.index(key)
                full = np.abs(save["dyfull"][im, idd] - z["dyfull"][d, m]).max() / np.abs(z["dyfull"][d, m]).max()
                print(f"  {name} mode {m} k={z['kmodes'][m]:.3g} {key}: raw state tangent {full:.2e} of its largest entry; "
                      f"dP: {np.abs(save['dpk4'][im, idd] / z['dpk4'][d, m] - 1).max():.2e}", flush=True)
        print(name, "%.1fs" % (time.time() - t), flush=True)


if __name__ == "__main__":
    main()
