"""Per-mode timeline of a team-kernel launch (library built with -DDEB_TEAM_TIMING): start, end, SM, slot -> us per
step of every mode, and how it depends on what else was running on the SM."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
lib = _cabi.default_library()
tab = helpers.load_tables("fiducial")
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ks = np.geomspace(1e-4, 10.0, nk)
dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=2048, power_idx=4)
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
os.environ["DEB_VARIANT"] = "team"
for _ in range(3):
    out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
buf = (ctypes.c_longlong * (4 * nk))()
lib.lib.deb_debug_team_mode_log(buf, nk)
L = np.array(list(buf), dtype=np.int64).reshape(nk, 4)
t0 = L[:, 0].min()
st, en, sm, slot = (L[:, 0] - t0) * 1e-3, (L[:, 1] - t0) * 1e-3, L[:, 2], L[:, 3]
ns = out["nsteps"][0]
us = (en - st) / ns
print(f"kernel_ms {out['kernel_ms']:.2f} (timing build) last end {en.max()*1e-3:.2f} ms")
order = np.argsort(-en)
print("latest finishing modes: idx k steps start_ms end_ms us/step sm slot")
for m in order[:8]:
    print(f"  {m:4d} {ks[m]:9.4g} {ns[m]:4d} {st[m]*1e-3:7.2f} {en[m]*1e-3:7.2f} {us[m]:6.1f} {sm[m]:4d} {slot[m]}")
long = ns > 300
print(f"modes with > 300 steps: {long.sum()}  us/step min {us[long].min():.1f} median {np.median(us[long]):.1f} max {us[long].max():.1f}")
first = st < 50.0
print(f"modes started in the first wave: {first.sum()}; later: {(~first).sum()}; latest start {st.max()*1e-3:.2f} ms")
# occupancy of each SM over time: number of resident modes while mode m runs (time-averaged)
occ = np.zeros(nk)
for m in range(nk):
    same = (sm == sm[m])
    ov = np.clip(np.minimum(en[same], en[m]) - np.maximum(st[same], st[m]), 0, None).sum() / (en[m] - st[m])
    occ[m] = ov
for lo, hi in ((0.9, 1.2), (1.2, 1.6), (1.6, 2.1)):
    sel = long & (occ >= lo) & (occ < hi)
    if sel.any():
        print(f"  long modes with mean SM occupancy in [{lo},{hi}): {sel.sum():3d}  median us/step {np.median(us[sel]):.1f}")
