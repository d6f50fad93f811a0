"""Hybrid launch (team CTAs for the longest modes + chain-lane warps for the rest) against the plain team launch on the
bench workload and neighbours; results are compared with the plain team launch (free-running tolerance)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
lib = _cabi.default_library()
ora = helpers.load_case("config2_full512")
for nk in [int(a) for a in sys.argv[1:]] or [296, 512, 768, 1024]:
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=2048, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    res = {}
    for name, env in (("team", {}), ("hybrid", {"DEB_HYBRID": "1"})):
        os.environ.pop("DEB_HYBRID", None); os.environ.update(env)
        ts = []
        for _ in range(8):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
            ts.append(out["kernel_ms"])
        res[name] = (min(ts[2:]), float(np.median(ts[2:])), ts[0], out)
    a, b = res["team"][3], res["hybrid"][3]
    rel = np.abs(b["pk"] / a["pk"] - 1)
    row = dict(modes=nk, team_ms=round(res["team"][0], 3), hybrid_ms=round(res["hybrid"][0], 3), hybrid_first_call_ms=round(res["hybrid"][2], 3),
               status_ok=bool(np.all(b["status"] == 0)), pk_vs_team_median=float(np.median(rel)), pk_vs_team_max=float(rel.max()))
    if nk == 512:
        ro = np.abs(b["pk"][0, :, 0] / ora["pk4"][:, 0] - 1)
        row.update(frac_within_1e5_of_oracle=float((ro < 1e-5).mean()), max_rel_vs_oracle=float(ro.max()))
    print(json.dumps(row), flush=True)
