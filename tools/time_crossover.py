"""Where the CTA-per-mode kernel stops paying: launch time against the number of modes for the automatic choice of the
one-warp kernels (DEB_VARIANT=warp / helper) and the team kernel, n = 265 and n = 72."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
for dm in ((31, 31, 31, 31, 5), (11, 11, 11, 8, 3)):
    lg, lp, lr, ln, nq = dm
    for nk in (512, 768, 1024, 1536, 2048, 4096):
        ks = np.geomspace(1e-4, 10.0, nk)
        dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
        ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
        res = {}
        for v in ("warp", "team"):
            os.environ["DEB_VARIANT"] = v
            best = 1e9
            for _ in range(3):
                out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
                best = min(best, out["kernel_ms"])
            res[v] = best
        print(f"n {lib.nvar(*dm):3d} nk {nk:5d} one-warp {res['warp']:7.2f} ms  team {res['team']:7.2f} ms  modes/s {nk/min(res.values())*1e3:8.0f}", flush=True)
