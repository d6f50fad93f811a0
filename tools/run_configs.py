"""Runs the BASELINE.json configurations that fit one GPU and prints a table."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
lib = _cabi.default_library()
T = {n: helpers.load_tables(n) for n in ("fiducial", "w0wa", "massless", "fiducial_nt1024")}
rows = []
def run(name, tabs, dm, nk, aout, rtol, kmin=1e-4, kmax=10.0, reps=3):
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(kmin, kmax, nk)
    dims = _cabi.make_dims(ncosmo=len(tabs), nk=nk, nout=len(aout), lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tabs[0].nth,
                           nnu=tabs[0].nnu, max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    sc = np.stack([t.scalars for t in tabs]); tb = np.stack([t.tables for t in tabs])
    best = 1e9
    for _ in range(reps):
        out = lib.evolve_host(dims, ctrl, sc, tb, ks, np.asarray(aout, dtype=float), want_pk=True)
        best = min(best, out["kernel_ms"])
    nm = len(tabs) * nk
    r = dict(config=name, n=lib.nvar(*dm), modes=nm, kernel_ms=round(best, 2), modes_per_s=round(nm / best * 1e3), steps=int(out["nsteps"].sum()),
             max_steps_mode=int(out["nsteps"].max()), ok=bool(np.all(out["status"] == 0)))
    rows.append(r); print(json.dumps(r), flush=True)
run("1: LCDM massless nu, lmax 16, 64 k, z=0", [T["massless"]], (16, 16, 16, 16, 3), 64, [1.0], 1e-4)
run("2: LCDM + massive nu, lmax 31 nq 5, 512 k, z=0", [T["fiducial"]], (31, 31, 31, 31, 5), 512, [1.0], 1e-4)
run("2': same at the CLASS-test shape z=99, k in [1e-5,10]", [T["fiducial"]], (31, 31, 31, 31, 5), 512, [0.01], 1e-4, kmin=1e-5)
run("3: w0wa + massive nu, 4096 k", [T["w0wa"]], (31, 31, 31, 31, 5), 4096, [1.0], 1e-4)
cyc = [T["fiducial"], T["w0wa"], T["massless"]]
run("4 (1/8 share): 128 cosmologies x 256 k (3 distinct tables cycled)", [cyc[i % 3] for i in range(128)], (31, 31, 31, 31, 5), 256, [1.0], 1e-4, reps=2)
run("reference benchmark shape: n=72, 256 k in [1e-5,10], z=99, rtol 1e-3", [T["fiducial"]], (11, 11, 11, 8, 3), 256, [0.01], 1e-3, kmin=1e-5)
run("reference pytest-benchmark workload (tests/test_perturbations.py:70-77): n=72, 256 k in [1e-5,1e2], z=99, rtol 1e-3, num_thermo=1024 tables", [T["fiducial_nt1024"]], (11, 11, 11, 8, 3), 256, [0.01], 1e-3, kmin=1e-5, kmax=1e2)
run("n=72 throughput: 8192 k", [T["fiducial"]], (11, 11, 11, 8, 3), 8192, [1.0], 1e-3)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "configs_r1.json"), "w"), indent=1)
