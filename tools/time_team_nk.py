"""Launch time of the CTA-per-mode kernel against the number of modes (same k range): shows where a launch stops being
bound by its slowest mode and starts being bound by the number of resident CTAs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.Library(sys.argv[1]) if len(sys.argv) > 1 else _cabi.default_library()
os.environ["DEB_VARIANT"] = "team"
for minb in ("2",):
    os.environ["DEB_TEAM_MINB"] = minb
    for nk in ((64, 296, 512) if len(sys.argv) > 1 else (64, 148, 256, 296, 320, 384, 444, 512, 592)):
        ks = np.geomspace(1e-4, 10.0, nk)
        dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
        ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
        best = 1e9
        for _ in range(4):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
            best = min(best, out["kernel_ms"])
        print(f"MINB {minb} nk {nk:4d} kernel_ms {best:7.2f} max steps {out['nsteps'].max()} us/step of the slowest mode {1e3*best/out['nsteps'].max():.1f} total steps {out['nsteps'].sum()}", flush=True)
