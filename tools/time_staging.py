"""A/B of the table staging (north_star: "tables staged into shared memory by TMA"): DEB_STAGE_TABLES=1 brings the three
RHS splines (24 KB) into shared memory with one cp.async.bulk per CTA; =0 reads them through the data cache.  Same kernel,
same occupancy, single-cosmology launches (the staged copy is per CTA)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
os.environ["DEB_VARIANT"] = "lane"
for dm, nk in (((11, 11, 11, 8, 3), 16384), ((16, 16, 16, 16, 3), 8192), ((31, 31, 31, 31, 5), 4096)):
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    res = {}
    for stage in ("0", "1"):
        os.environ["DEB_STAGE_TABLES"] = stage
        ts = []
        for _ in range(5):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
            ts.append(out["kernel_ms"])
        res[stage] = (min(ts), float(np.median(ts)), out["pk"].copy(), int(out["nsteps"].sum()))
    same = bool(np.array_equal(res["0"][2], res["1"][2]))
    print(json.dumps(dict(n=lib.nvar(*dm), modes=nk, ldg_ms_best=round(res["0"][0], 3), ldg_ms_median=round(res["0"][1], 3), staged_ms_best=round(res["1"][0], 3),
                          staged_ms_median=round(res["1"][1], 3), identical_results=same, steps=res["0"][3])), flush=True)
