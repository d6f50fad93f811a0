"""Generates tests/golden/tangent_<case>.npz: forward-tangent vectors of the hot path from the complex-step
oracle (oracle/discoeb_tangent.py), together with the packed inputs and tangent seeds the CUDA path is fed.

Seeds: a cosmological parameter is shifted by +-1e-4 (relative) in the CPU table producer
(oracle/background.py) and every scalar and spline array is differenced (SURVEY.md section 8d, config 5:
"table tangents from central differences of the CPU table producer").  The parity tests do not depend on how
good that derivative is: oracle and kernel consume the same seeds.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
import oracle.background as B  # noqa: E402
import oracle.discoeb_tangent as T  # noqa: E402
from discoeb_b200 import _pack  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def param_tangent(key, rel=1e-4, **over):
    """(param, dparam) with dparam = d param / d key by central differences of the table producer.  Keys that
    only enter the post-processing (A_s, n_s, k_p) get their exact unit seed."""
    base = B.fiducial_param(**over)
    p0 = B.evolve_background(dict(base))
    if key in ("A_s", "n_s", "k_p"):
        dp = {k: 0.0 for k in T.SCALAR_KEYS if k in p0}
        dp[key] = 1.0
        for k in T.SPLINE_KEYS:
            s = B.Spline.__new__(B.Spline)
            s.x, s.y, s.S = np.zeros_like(p0[k].x), np.zeros_like(p0[k].y), np.zeros_like(p0[k].S)
            dp[k] = s
        return p0, dp
    h = rel * abs(base[key]) if base[key] != 0 else rel
    pp = B.evolve_background(dict(base, **{key: base[key] + h}))
    pm = B.evolve_background(dict(base, **{key: base[key] - h}))
    dp = {}
    for k in T.SCALAR_KEYS:
        if k in p0:
            dp[k] = (float(pp[k]) - float(pm[k])) / (2 * h)
    for k in T.SPLINE_KEYS:
        s = B.Spline.__new__(B.Spline)
        s.x, s.y, s.S = (pp[k].x - pm[k].x) / (2 * h), (pp[k].y - pm[k].y) / (2 * h), (pp[k].S - pm[k].S) / (2 * h)
        dp[k] = s
    return p0, dp


def pack_tangent(dp):
    """dparam -> (d_scalars[NSCAL], d_tables[tl]) in the C-ABI layout."""
    dscal = np.zeros(_pack.NSCAL)
    for i, k in enumerate(_pack.SCALAR_KEYS):
        dscal[i] = dp.get(k, 0.0)
    dtab = np.concatenate([np.concatenate([dp[k].x, dp[k].y, dp[k].S]) for k in _pack.SPLINE_KEYS])
    return dscal, dtab


CASES = {
    # name: (cosmology overrides, dims, kmodes, aexp_out, rtol, directions)
    "default_n72": ({}, (11, 11, 11, 8, 3), [1e-3, 0.05, 0.5], [0.1, 1.0], 1e-4, ("Omegam", "H0", "w_DE_0", "n_s")),
    "w0wa_n43": (dict(w_DE_0=-0.9, w_DE_a=0.1), (5, 4, 6, 3, 4), [2e-3, 0.2], [0.5], 1e-3, ("w_DE_a", "Omegab")),
    "fisher_n265": ({}, (31, 31, 31, 31, 5), [0.01, 0.3], [0.5, 1.0], 1e-4, ("Omegam",)),
    "fisher_n265x2": ({}, (31, 31, 31, 31, 5), [0.02], [0.5, 1.0], 1e-4, ("Omegab", "H0")),      # seeds for the full-size property test
    # BASELINE config 5 (Fisher Jacobian: n=265, a_out = [0.5, 1], 7 directions): every 32nd mode of the 512-mode grid,
    # the ~600-step k=10 mode included
    "config5_n265": ({}, (31, 31, 31, 31, 5), list(np.geomspace(1e-4, 10.0, 512)[31::32]), [0.5, 1.0], 1e-4,
                     ("Omegam", "Omegab", "w_DE_0", "w_DE_a", "H0", "n_s", "A_s")),
    # wavenumbers that move with the parameter (k = const x h, nb_discoeb_rsd_eyes_plot.ipynb cell 5): d k / d H0 = k / H0;
    # the second direction moves k alone (all other seeds zero)
    "kscaled_n72": ({}, (11, 11, 11, 8, 3), [2e-3, 0.1], [0.3, 1.0], 1e-4, ("H0", "n_s")),
}
DKMODES = {"kscaled_n72": lambda ks, p: np.stack([ks / p["H0"], ks])}
KEEP = ("y", "dy", "yfull", "dyfull", "pk4", "dpk4", "tau_out", "dtau_out", "tau_start", "dtau_start", "y0", "dy0",
        "rp_tnext", "rp_dtnext", "rp_keep", "nsteps", "naccept")


def main():
    only = sys.argv[1:]
    for name, (over, dims, ks, aout, rtol, dirs) in CASES.items():
        if only and name not in only:
            continue
        t = time.time()
        ks = np.asarray(ks, dtype=np.float64)
        outs, dscal, dtab = [], [], []
        replay = None
        dks = None
        for idir, key in enumerate(dirs):
            p, dp = param_tangent(key, **over)
            if name in DKMODES:
                dks = DKMODES[name](ks, p)
                if key == "n_s":
                    dp["n_s"] = 0.0            # a pure wavenumber direction
            # direction 0 runs the controller; the others follow its decisions (the real part of a complex run carries
            # direction-dependent round-off, and the controller is chaotic under round-off -- DESIGN.md "Parity")
            out = T.evolve_perturbations_jvp(param=p, dparam=dp, aexp_out=aout, kmodes=ks, rtol=rtol, atol=rtol,
                                             lmaxg=dims[0], lmaxgp=dims[1], lmaxr=dims[2], lmaxnu=dims[3], nqmax=dims[4],
                                             max_steps=4096, replay=replay, dkmodes=None if dks is None else dks[idir])
            if replay is None:
                replay = (out["rp_keep"], out["rp_fac"], out["nsteps"])
            outs.append(out)
            a, b = pack_tangent(dp)
            dscal.append(a)
            dtab.append(b)
        scal, tab, nth, nnu = _pack.pack_param(p)
        for o in outs[1:]:      # every direction repeats the same primal run
            assert np.array_equal(o["rp_keep"], outs[0]["rp_keep"]) and np.allclose(o["rp_tnext"], outs[0]["rp_tnext"], rtol=1e-9, atol=0)
        save = dict(scalars=scal, tables=tab, nth=nth, nnu=nnu, d_scalars=np.stack(dscal), d_tables=np.stack(dtab),
                    kmodes=ks, aexp_out=np.asarray(aout, dtype=np.float64), dims=np.array(dims), rtol=rtol,
                    directions=np.array(dirs))
        if dks is not None:
            save["d_kmodes"] = dks
        for f in KEEP:
            if f.startswith("d") or f == "rp_dtnext":
                save[f] = np.stack([o[f] for o in outs])
            else:
                save[f] = outs[0][f]
        np.savez_compressed(os.path.join(GOLD, f"tangent_{name}.npz"), **save)
        print(name, "steps", outs[0]["nsteps"], "%.1fs" % (time.time() - t), flush=True)


if __name__ == "__main__":
    main()
