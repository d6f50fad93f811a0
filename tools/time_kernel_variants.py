"""Times the bench workload (and the default-hierarchy one) under each kernel variant of the library
(DEB_VARIANT / DEB_TEAM / DEB_TEAM_MINB are read at launch time) and checks that the variants agree."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
VARIANTS = [("helper", {}), ("warp", {}), ("team", {"DEB_TEAM": "4", "DEB_TEAM_MINB": "4"}), ("team", {"DEB_TEAM": "4", "DEB_TEAM_MINB": "3"}),
            ("team", {"DEB_TEAM": "4", "DEB_TEAM_MINB": "2"}), ("team", {"DEB_TEAM": "8"})]
if len(sys.argv) > 1:
    VARIANTS = [v for v in VARIANTS if v[0] in sys.argv[1:]]
for (dm, nk, rtol, kmin, kmax, aout) in [((31, 31, 31, 31, 5), 512, 1e-4, 1e-4, 10.0, 1.0), ((11, 11, 11, 8, 3), 512, 1e-4, 1e-4, 10.0, 1.0),
                                         ((16, 16, 16, 16, 3), 64, 1e-4, 1e-4, 10.0, 1.0), ((31, 31, 31, 31, 5), 512, 1e-4, 1e-5, 10.0, 0.01)]:
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(kmin, kmax, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    ref = None
    for name, env in VARIANTS:
        for k in ("DEB_VARIANT", "DEB_TEAM", "DEB_TEAM_MINB"):
            os.environ.pop(k, None)
        os.environ["DEB_VARIANT"] = name
        os.environ.update(env)
        best = 1e9
        for _ in range(4):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([aout]), want_pk=True)
            best = min(best, out["kernel_ms"])
        if name == "helper":
            ref = out
        same = "" if ref is None else f" bitwise==helper {np.array_equal(out['y'], ref['y'])} maxrel {np.nanmax(np.abs(out['pk'] / ref['pk'] - 1)):.1e}"
        print(f"n={lib.nvar(*dm):3d} nk={nk:5d} a_out={aout} {name:6s} {str(env):48s} kernel_ms {best:8.2f} modes/s {nk/best*1e3:9.0f} steps {out['nsteps'].sum()} max {out['nsteps'].max()} status {out['status'].max()}{same}", flush=True)
