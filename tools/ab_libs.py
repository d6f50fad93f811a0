"""A/B of two builds of the library (argv[1], argv[2] = paths of the .so files) on the one-warp, lane, team and tangent kernels."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
tab = helpers.load_tables("fiducial")
libs = {os.path.basename(p): _cabi.Library(p) for p in sys.argv[1:3]}
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
cases = [("n72 16384 warp", (11, 11, 11, 8, 3), 16384, "warp"), ("n111 8192 warp", (16, 16, 16, 16, 3), 8192, "warp"), ("n265 4096 lane", (31, 31, 31, 31, 5), 4096, "lane"),
         ("n265 4096 warp", (31, 31, 31, 31, 5), 4096, "warp"), ("n265 512 team", (31, 31, 31, 31, 5), 512, "team")]
for name, dm, nk, var in cases:
    os.environ["DEB_VARIANT"] = var
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
    row, outs = dict(case=name), {}
    for ln_, lib in libs.items():
        ts = []
        for _ in range(4):
            out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
            ts.append(out["kernel_ms"])
        row[ln_] = round(min(ts[1:]), 3); outs[ln_] = out
    a, b = list(outs.values())
    row["bit_identical"] = bool(np.array_equal(a["pk"], b["pk"]) and np.array_equal(a["nsteps"], b["nsteps"]))
    print(json.dumps(row), flush=True)
os.environ.pop("DEB_VARIANT", None)
z = np.load(os.path.join(ROOT, "tests", "golden", "fisher_seeds.npz"))
ks = np.geomspace(1e-4, 10.0, 512)
for ntan in (1, 7):
    dims = _cabi.make_dims(ncosmo=1, nk=512, nout=2, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=int(z["nth"]), nnu=int(z["nnu"]), max_steps=4096, power_idx=4, ntan=ntan)
    row, outs = dict(case=f"tangent 512 x {ntan}"), {}
    for ln_, lib in libs.items():
        out = lib.evolve_tangent_host(dims, ctrl, z["scalars"][None], z["tables"][None], ks, np.array([0.5, 1.0]), z["d_scalars"][:ntan, None], z["d_tables"][:ntan, None], want_pk=True)
        row[ln_] = round(out["kernel_ms"], 2); outs[ln_] = out
    a, b = list(outs.values())
    row["bit_identical"] = bool(np.array_equal(a["dy"], b["dy"]))
    print(json.dumps(row), flush=True)
