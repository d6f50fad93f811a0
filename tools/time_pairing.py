"""What does a second CTA on the SM cost the team kernel, and is it the shared instruction cache?  296 copies of ONE
mode run in lock-step (both CTAs of an SM execute the same code at the same time); two different modes do not."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
os.environ["DEB_VARIANT"] = "team"
def run(ks, label):
    nk = len(ks)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    best = 1e9
    for _ in range(4):
        out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
        best = min(best, out["kernel_ms"])
    print(f"{label:58s} nk {nk:4d} kernel_ms {best:7.2f} max steps {out['nsteps'].max()} us/step {1e3*best/out['nsteps'].max():.1f}", flush=True)
run(np.full(148, 0.145), "148 copies of k=0.145 (one CTA per SM)")
run(np.full(296, 0.145), "296 copies of k=0.145 (two CTAs per SM, lock-step)")
run(np.concatenate([np.full(148, 0.145), np.full(148, 0.1449)]), "148 x k=0.145 + 148 x k=0.1449 (two CTAs, near lock-step)")
run(np.concatenate([np.full(148, 0.145), np.full(148, 0.09)]), "148 x k=0.145 + 148 x k=0.09 (two CTAs, different phases)")
run(np.concatenate([np.full(148, 0.145), np.full(148, 3.0)]), "148 x k=0.145 + 148 x k=3 (two CTAs, different phases)")
