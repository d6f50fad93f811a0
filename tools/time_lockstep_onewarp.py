"""Is the one-warp throughput kernel bound by instruction supply?  N copies of one mode keep all warps of an SM in
lock-step (same code at the same time); distinct modes do not.  Compares ns per (mode, step)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
os.environ["DEB_VARIANT"] = "warp"
def run(ks, label, dm=(31, 31, 31, 31, 5)):
    nk = len(ks)
    lg, lp, lr, ln, nq = dm
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    best = 1e9
    for _ in range(3):
        out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
        best = min(best, out["kernel_ms"])
    tot = int(out["nsteps"].sum())
    print(f"{label:52s} nk {nk:5d} kernel_ms {best:8.2f} total steps {tot:8d} ns per step {1e6*best/tot:7.1f} modes/s {nk/best*1e3:8.0f}", flush=True)
run(np.geomspace(1e-4, 10.0, 4096), "4096 distinct modes (config-3 k grid)")
run(np.full(4096, 0.145), "4096 copies of k=0.145")
run(np.full(1184, 0.145), "1184 copies of k=0.145 (one wave, 8 per SM)")
run(np.geomspace(0.05, 10.0, 1184), "1184 distinct modes k in [0.05,10] (one wave)")
run(np.geomspace(1e-4, 10.0, 8192), "8192 distinct modes, n=72", dm=(11, 11, 11, 8, 3))
run(np.full(8192, 0.145), "8192 copies of k=0.145, n=72", dm=(11, 11, 11, 8, 3))
