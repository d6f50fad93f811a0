"""Times the bench workload with alternative builds of the library (ablation runs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
for path in sys.argv[1:]:
    lib = _cabi.Library(path)
    for (dm, nk, rtol) in [((11, 11, 11, 8, 3), 512, 1e-3), ((31, 31, 31, 31, 5), 512, 1e-4), ((31, 31, 31, 31, 5), 4096, 1e-4), ((11, 11, 11, 8, 3), 8192, 1e-3)]:
        lg, lp, lr, ln, nq = dm
        ks = np.geomspace(1e-4, 10.0, nk)
        dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096)
        ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
        best = 1e9
        for _ in range(3):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]))
            best = min(best, out["kernel_ms"])
        print(f"{os.path.basename(path):28s} n={lib.nvar(*dm):3d} nk={nk:5d} kernel_ms {best:8.2f}  modes/s {nk/best*1e3:9.0f}  steps/s {out['nsteps'].sum()/best*1e3:.3e}", flush=True)
