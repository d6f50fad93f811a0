"""Runs the bench workload (config 2: 512 modes, n=265) a few times through the host C-ABI; the target of ncu captures."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
lib = _cabi.default_library()
tab = helpers.load_tables("fiducial")
ks = np.geomspace(1e-4, 10.0, nk)
dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=2048, power_idx=4)
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
for _ in range(reps):
    out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
    print("kernel_ms", out["kernel_ms"], "steps", int(out["nsteps"].sum()), "status", np.unique(out["status"]))
