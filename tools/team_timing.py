"""Phase timing of warp 0 of the CTA-per-mode kernel (library built with EXTRA=-DDEB_TEAM_TIMING): cycles per step
spent in each phase by the largest-k mode of the bench workload."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from discoeb_b200 import _cabi
lib = _cabi.Library(os.path.join(ROOT, 'disco-eb_b200', 'csrc', '_obj', 'libdeb_timing.so')) if os.path.exists(os.path.join(ROOT, 'disco-eb_b200', 'csrc', '_obj', 'libdeb_timing.so')) else _cabi.default_library()
tab = helpers.load_tables("fiducial")
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kmax = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
aout = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
ks = np.geomspace(1e-4, kmax, nk)
dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=2048, power_idx=4)
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
os.environ["DEB_VARIANT"] = "team"
buf = (ctypes.c_longlong * 32)()
skip = int(os.environ.get("DEB_TEAM_SKIP", "0"))
lib.lib.deb_debug_team_skip(skip)       # knock phases out (wrong results on purpose): what does a phase cost the step?
if skip:
    dims.max_steps = 300
lib.lib.deb_debug_team_timing(buf, 1)
out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([aout]), want_pk=True)
lib.lib.deb_debug_team_timing(buf, 0)
t = np.array(list(buf), dtype=float)
steps = t[15]
names = ["J: rows/factors (own work)", "J: wait barrier", "J: block inverses + Woodbury", "stage: own elements", "stage: wait B1", "stage: metric+head rows (+x0)",
         "stage: sweeps by the last warp (as seen by warp 0)", "stage: head solve + carries (warp 0)", "stage: wait B4", "step end: norm, controller, copy"]
print(f"skip mask {skip} kmax {kmax} a_out {aout} kernel_ms {out['kernel_ms']:.2f} steps of the largest-k mode {int(steps)} cycles/step {t[:10].sum()/steps:.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:42s} {t[i]/steps:9.0f} cycles/step  {100*t[i]/t[:10].sum():5.1f} %")
for i, nm in zip(range(11, 15), ["solve: wait for the swept tails + l=2 rows", "solve: gather + block multiply", "solve: Woodbury sums + write-back", "solve: wait for forward recurrences + carry chain"]):
    print(f"    {nm:40s} {t[i]/steps/8:9.0f} cycles/stage")
print("  whole stage by stage number (cycles per step): " + "  ".join(f"st{st}: {t[16 + st] / steps:.0f}" for st in range(1, 9)))
