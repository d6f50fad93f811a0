"""Times the GPU table producer (deb_background_host_f64) for batches of BASELINE config-4 cosmologies and, beside it,
the end-to-end chain parameters -> tables -> 256 k-modes per cosmology (config 4's per-GPU share)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from discoeb_b200 import _cabi
from discoeb_b200.background import config4_draws, pack_background_input
lib = _cabi.default_library()
base = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09, n_s=0.96822,
            H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06)
rows = []
for nc in [int(a) for a in sys.argv[1:]] or [1, 16, 128, 296, 1024]:
    bg_in = np.stack([pack_background_input({**base, **d}) for d in config4_draws(nc)])
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); scal, tab, ms = lib.background_host(bg_in, 256); wall = 1e3 * (time.perf_counter() - t)
        best = min(best, ms)
    r = dict(what="tables", ncosmo=nc, kernel_ms=round(best, 3), host_call_ms=round(wall, 2), cosmologies_per_s=round(nc / best * 1e3))
    rows.append(r); print(json.dumps(r), flush=True)
# config 4, one GPU's share: 128 distinct cosmologies x 256 k, tables produced on the device first
nc, nk = 128, 256
bg_in = np.stack([pack_background_input({**base, **d}) for d in config4_draws(1024)[:nc]])
scal, tab, ms_bg = lib.background_host(bg_in, 256)
ks = np.geomspace(1e-4, 10.0, nk)
dims = _cabi.make_dims(ncosmo=nc, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=256, nnu=512, max_steps=4096, power_idx=4)
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
best = 1e9
for _ in range(2):
    out = lib.evolve_host(dims, ctrl, scal, tab, ks, np.array([1.0]), want_pk=True)
    best = min(best, out["kernel_ms"])
r = dict(what="config4 share: 128 distinct cosmologies x 256 k (n=265)", tables_ms=round(ms_bg, 2), solve_ms=round(best, 2), modes=nc * nk,
         modes_per_s_solve=round(nc * nk / best * 1e3), modes_per_s_end_to_end=round(nc * nk / (best + ms_bg) * 1e3),
         status_ok=bool(np.all(out["status"] == 0)), steps=int(out["nsteps"].sum()))
rows.append(r); print(json.dumps(r), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "background_timing.json"), "w"), indent=1)
