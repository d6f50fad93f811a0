"""Two teams per CTA with their serial warps in lock-step (k_evolve_duo) against two CTAs of one team (k_evolve_team) on
the bench workload (512 k, n=265) and neighbours.  The rendezvous moves no data: results must be bit-identical."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
variants = [("solo", {"DEB_DUO": "0"}), ("duo_partnered", {"DEB_DUO": "1", "DEB_DUO_SOLO": "0"})] + [(f"duo_ls{e}", {"DEB_DUO": "1", "DEB_DUO_LOCKSTEP": str(e)}) for e in (0, 2, 1)]
for nk in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=2048, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    row, base = dict(modes=nk), None
    for name, env in variants:
        for k_ in ("DEB_DUO", "DEB_DUO_LOCKSTEP", "DEB_DUO_SOLO"): os.environ.pop(k_, None)
        os.environ.update(env)
        ts = []
        for _ in range(8):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
            ts.append(out["kernel_ms"])
        row[name + "_ms"] = round(min(ts[2:]), 3)
        if base is None: base = out
        else: row[name + "_bit_identical"] = bool(np.array_equal(out["pk"], base["pk"]) and np.array_equal(out["nsteps"], base["nsteps"]) and np.all(out["status"] == 0))
        print(json.dumps(row), flush=True)
