"""First-light check on a B200: CUDA library vs the CPU emulation of the same source vs the oracle."""
import os, sys, time, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
import oracle.background as B
import oracle.discoeb_oracle as O
from discoeb_b200 import _cabi, _pack

emu_so = os.path.join(ROOT, "tests/emu/_build/libdeb_emu.so")
if not os.path.exists(emu_so):
    os.makedirs(os.path.dirname(emu_so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-o", emu_so,
                           os.path.join(ROOT, "tests/emu/deb_emu.cpp")])
emu = _cabi.Library(emu_so, prefix="emu_")
gpu = _cabi.Library()
print("devices", gpu.lib.deb_device_count())
p = B.evolve_background(B.fiducial_param())
scal, tab, nth, nnu = _pack.pack_param(p)
res = {}
for (lg, lp, lr, ln, nq) in [(11, 11, 11, 8, 3), (16, 16, 16, 16, 3), (31, 31, 31, 31, 5)]:
    d = O.Dims(lg, lp, lr, ln, nq)
    ks = np.geomspace(1e-4, 10, 8); aout = np.array([0.01, 1.0])
    dims = _cabi.make_dims(ncosmo=1, nk=len(ks), nout=len(aout), lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq,
                           nth=nth, nnu=nnu, max_steps=4096)
    ts_g, y0_g = gpu.debug_ics(dims, scal, tab, ks, aout)
    ts_e, y0_e = emu.debug_ics(dims, scal, tab, ks, aout)
    print("n", d.n, "ICs gpu-vs-emu", np.abs(ts_g / ts_e - 1).max(), (np.abs(y0_g - y0_e) / (np.abs(y0_e) + 1e-300)).max())
    rng = np.random.default_rng(0)
    t0 = np.array([1.0, 5.0, 20., 50., 100., 200., 300., 1000.])
    y = rng.normal(size=(8, d.n)); y[:, 0] = p['a_of_tau_spline'].evaluate(t0)
    t1 = t0 * (1 + np.array([0.2, 0.1, 0.05, 0.02, 0.01, 0.01, 0.003, 0.1]))
    y1g, eg = gpu.debug_step(dims, scal, tab, ks, t0, t1, y)
    y1e, ee = emu.debug_step(dims, scal, tab, ks, t0, t1, y)
    y1o, eo = O.rodas5_step(t0, t1, y, p, ks, d)
    sc = np.abs(y1o).max(axis=1, keepdims=True)
    print("   step gpu-vs-emu", (np.abs(y1g - y1e) / sc).max(), " gpu-vs-oracle", (np.abs(y1g - y1o) / sc).max())
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    t = time.time(); og = gpu.evolve_host(dims, ctrl, scal[None], tab[None], ks, aout); tg = time.time() - t
    oe = emu.evolve_host(dims, ctrl, scal[None], tab[None], ks, aout)
    print("   evolve status", og['status'][0], "nsteps gpu", og['nsteps'][0], "emu", oe['nsteps'][0], "kernel_ms", og['kernel_ms'], "wall", tg)
    sc = np.abs(oe['y']).max(axis=(0, 1, 2))
    print("   evolve gpu-vs-emu per-field max|diff|/fieldmax", (np.abs(og['y'] - oe['y']).max(axis=(0, 1, 2)) / sc).max())

# config 2b: 512 modes, n=265
for (name, dm, nk, rtol) in [("default n=72", (11, 11, 11, 8, 3), 512, 1e-3), ("config2b n=265", (31, 31, 31, 31, 5), 512, 1e-4),
                             ("n=265 x2048", (31, 31, 31, 31, 5), 2048, 1e-4)]:
    lg, lp, lr, ln, nq = dm
    ks = np.geomspace(1e-4, 10, nk); aout = np.array([1.0])
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=nth, nnu=nnu, max_steps=4096)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    for rep in range(3):
        t = time.time(); og = gpu.evolve_host(dims, ctrl, scal[None], tab[None], ks, aout); tg = time.time() - t
    t = time.time(); oe = emu.evolve_host(dims, ctrl, scal[None], tab[None], ks, aout); te = time.time() - t
    st = int(og['nsteps'].sum())
    print(name, "nk", nk, "kernel_ms", og['kernel_ms'], "wall_ms", tg * 1e3, "status", np.unique(og['status']), "steps", st, "max", og['nsteps'].max(),
          "modes/s", nk / (og['kernel_ms'] * 1e-3), "steps/s", st / (og['kernel_ms'] * 1e-3), "| emu(cpu %d thr) s" % os.cpu_count(), te)
    d4 = np.abs(og['y'][0, :, 0, 4] / oe['y'][0, :, 0, 4] - 1)
    print("   delta_m gpu-vs-emu max rel", d4.max(), "median", np.median(d4), "nsteps equal frac", np.mean(og['nsteps'] == oe['nsteps']))
    res[name] = dict(kernel_ms=og['kernel_ms'], steps=st)
print("fp64 peak", gpu.fp64_peak_tflops())
json.dump(res, open(os.path.join(ROOT, "gpurun_out/first_light.json"), "w"))
