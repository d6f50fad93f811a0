"""At n = 72 and rtol 1e-4: how far the kernel variants (lane, cyclic one-warp, team) are from each other and from a
tight-tolerance run on a deep multi-cosmology launch -- the spread that sets the bar of
tests/test_gpu_parity.py::test_deep_launch_of_small_hierarchies_takes_the_lane_kernel (run under gpurun)."""
import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
lib = _cabi.default_library()
tabs = [helpers.load_tables("fiducial"), helpers.load_tables("w0wa"), helpers.load_tables("fiducial")]
sc = np.stack([t.scalars for t in tabs]); tb = np.stack([t.tables for t in tabs])
nk = 600
dims = _cabi.make_dims(ncosmo=3, nk=nk, nout=2, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tabs[0].nth, nnu=tabs[0].nnu, max_steps=8192, power_idx=4)
ks = np.geomspace(1e-4, 5.0, nk)
res = {}
for name, var, rt in (("lane", "lane", 1e-4), ("warp", "warp", 1e-4), ("team", "team", 1e-4), ("truth", "warp", 1e-8)):
    os.environ["DEB_VARIANT"] = var
    res[name] = lib.evolve_host(dims, _cabi.make_ctrl(rtol=rt, atol=rt), sc, tb, ks, np.array([0.3, 1.0]), want_pk=True)
t = res["truth"]["pk"]
for name in ("lane", "warp", "team"):
    r = np.abs(res[name]["pk"] / t - 1)
    i = np.unravel_index(np.argmax(r), r.shape)
    print(name, "vs truth: median", np.median(r), "max", r.max(), "at", i, "k", ks[i[1]], "steps", res[name]["nsteps"][i[0], i[1]], "frac>1e-2", (r > 1e-2).mean(), "frac>1e-3", (r > 1e-3).mean())
r = np.abs(res["lane"]["pk"] / res["warp"]["pk"] - 1); print("lane vs warp max", r.max(), "same steps", np.mean(res["lane"]["nsteps"] == res["warp"]["nsteps"]))
r = np.abs(res["team"]["pk"] / res["warp"]["pk"] - 1); print("team vs warp max", r.max(), "same steps", np.mean(res["team"]["nsteps"] == res["warp"]["nsteps"]))
