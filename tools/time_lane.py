"""Times the chain-lane kernel (DEB_VARIANT=lane) against the default kernel choice at several launch sizes (n = 265
and n = 72), and reports the deviation between the two (free-running solves, so O(rtol) on long modes)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
sizes = [int(a) for a in sys.argv[1:]] or [512, 4096, 16384]
for dm in [(31, 31, 31, 31, 5), (11, 11, 11, 8, 3)]:
    lg, lp, lr, ln, nq = dm
    for nk in sizes:
        ks = np.geomspace(1e-4, 10.0, nk)
        dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
        ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
        res = {}
        for name in ("default", "lane"):
            os.environ.pop("DEB_VARIANT", None)
            if name != "default":
                os.environ["DEB_VARIANT"] = name
            best = 1e9
            for _ in range(3):
                out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
                best = min(best, out["kernel_ms"])
            res[name] = out
            steps = int(out["nsteps"].sum())
            print(f"n={lib.nvar(*dm):3d} nk={nk:6d} {name:8s} kernel_ms {best:9.2f} modes/s {nk/best*1e3:9.0f} steps {steps} us/step/SM {best*1e3*148/steps:7.2f} "
                  f"max_steps_mode {out['nsteps'].max()} status {np.unique(out['status'])}", flush=True)
        d = np.abs(res["lane"]["pk"] / res["default"]["pk"] - 1)
        print(f"    lane vs default P(k): median {np.median(d):.2e} max {d.max():.2e} same step counts {np.mean(res['lane']['nsteps'] == res['default']['nsteps']):.2f}", flush=True)
