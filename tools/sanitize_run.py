"""Tiny workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
lib = _cabi.default_library()
tab = helpers.load_tables("fiducial")
for dm, nk in (((11, 11, 11, 8, 3), 6), ((31, 31, 31, 31, 5), 4), ((31, 31, 31, 31, 5), 21)):
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(1e-3, 0.05, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=2, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=60, power_idx=4)
    out = lib.evolve_host(dims, _cabi.make_ctrl(rtol=1e-3, atol=1e-3), tab.scalars[None], tab.tables[None], ks, np.array([0.001, 0.002]), want_pk=True)
    print("n", lib.nvar(*dm), "status", out["status"][0], "steps", out["nsteps"][0])
if os.environ.get("DEB_SANITIZE_PRIMAL_ONLY"):
    sys.exit(0)
# tangent kernel (one direction) and batched shared-step kernel (one CTA; a 2-CTA cluster)
z = np.load(os.path.join(ROOT, "tests", "golden", "fisher_seeds.npz"))
ks = np.geomspace(1e-3, 0.05, 3)
dims = _cabi.make_dims(ncosmo=1, nk=3, nout=2, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=int(z["nth"]), nnu=int(z["nnu"]), max_steps=40,
                       power_idx=4, ntan=2)
out = lib.evolve_tangent_host(dims, _cabi.make_ctrl(rtol=1e-3, atol=1e-3), z["scalars"][None], z["tables"][None], ks, np.array([0.001, 0.002]),
                              z["d_scalars"][:2, None], z["d_tables"][:2, None], want_pk=True, d_kmodes=np.stack([ks, 0 * ks]))
print("tangent status", out["status"][0], "steps", out["nsteps"][0], "finite", bool(np.all(np.isfinite(out["dy"]))))
for dm, nk, B in (((11, 11, 11, 8, 3), 8, 4), ((31, 31, 31, 31, 5), 16, 16)):
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(1e-3, 0.05, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=30,
                           batch_size=B)
    out = lib.evolve_host(dims, _cabi.make_ctrl(rtol=1e-3, atol=1e-3), tab.scalars[None], tab.tables[None], ks, np.array([0.001]))
    print("batched n", lib.nvar(*dm), "B", B, "status", out["status"][0], "steps", out["nsteps"][0])
