"""Where the team (two-team CTA) kernel hands over to the chain-lane kernel: device ms of both over the launch size."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
lib = _cabi.default_library()
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
for cos, dm in (("fiducial", (31, 31, 31, 31, 5)), ("fiducial", (11, 11, 11, 8, 3)), ("massless", (16, 16, 16, 16, 3))):
    tab = helpers.load_tables(cos)
    lg, lp, lr, ln, nq = dm
    for nk in [int(a) for a in sys.argv[1:]] or [444, 592, 740, 888, 1036, 1184, 1480, 1776, 2368]:
        ks = np.geomspace(1e-4, 10.0, nk)
        dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
        row = dict(n=lib.nvar(*dm), nk=nk, per_sm=round(nk / 148, 1))
        for name, env in (("default", {}), ("team", {"DEB_VARIANT": "team"}), ("lane", {"DEB_VARIANT": "lane"})):
            for k_ in ("DEB_VARIANT",): os.environ.pop(k_, None)
            os.environ.update(env)
            ts = []
            for _ in range(5):
                out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
                ts.append(out["kernel_ms"])
            row[name] = round(min(ts[2:]), 2)
        print(json.dumps(row), flush=True)
