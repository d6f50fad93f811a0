"""Oracle vectors on BASELINE.json's own configurations -> tests/golden/oracle_*.npz.

  config2_full512     all 512 modes of the bench workload (config 2, z = 0): P(k), 20 fields, step counts -- what the
                      `parity` block of the bench line and the full-size GPU test compare with
  config2_grid64_n265 every 8th mode of that grid WITH the step traces (replay parity, 64 modes, k up to 10/Mpc)
  config3_grid32_n265 w0wa + massive nu (config 3): every 128th mode of the 4096-mode grid, with traces
  config4_n265        3 numpy.random.default_rng(0) cosmologies x 16 k (config 4), tables = the REFERENCE's own
                      evolve_background output for those draws (tests/golden/reference_background.npz), stored alongside
  class_grid512       the reference's acceptance shape (z = 99, k in [1e-5, 10], 512 modes, rtol 1e-4): P_bc(k)
  converge_n265       32 modes of config 2 at rtol = atol = 1e-4, 1e-5, 1e-7 and the 1e-8 "truth" (convergence test)

    python tools/make_golden_baseline.py [name ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import helpers  # noqa: E402
import oracle.discoeb_oracle as O  # noqa: E402
from make_golden import run_case  # noqa: E402
from discoeb_b200._pack import SPLINE_KEYS  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
D265 = (31, 31, 31, 31, 5)


def outputs_only(p, dims, ks, aout, rtol, max_steps=4096, chunk=64):
    lg, lp, lr, ln, nq = dims
    y, k, _, info = O.evolve_perturbations(param=p, aexp_out=aout, kmin=0, kmax=0, num_k=len(ks), kmodes=ks, lmaxg=lg, lmaxgp=lp, lmaxr=lr,
                                           lmaxnu=ln, nqmax=nq, rtol=rtol, atol=rtol, max_steps=max_steps, return_info=True, chunk=chunk)
    return dict(kmodes=ks, aexp_out=np.asarray(aout, dtype=float), y=y, pk4=O.get_power(k=ks[:, None], y=y, idx=4, param=p),
                pk6=O.get_power(k=ks[:, None], y=y, idx=6, param=p), nsteps=info["nsteps"].astype(np.int32),
                naccept=info["naccept"].astype(np.int32), dims=np.array(dims), rtol=rtol)


def reference_tables(name):
    """Packed (scalars, tables) of a cosmology of tests/golden/reference_background.npz + the oracle param dict."""
    z = np.load(os.path.join(GOLD, "reference_background.npz"))
    scal = np.zeros(24)
    scal[:16] = z[f"{name}_scalars"]
    inp = z[f"{name}_in"]
    scal[16], scal[17], scal[18] = inp[12], inp[13], inp[14]
    tab = np.concatenate([np.concatenate([z[f"{name}_{k}_x"], z[f"{name}_{k}_y"], z[f"{name}_{k}_S"]]) for k in SPLINE_KEYS])
    return helpers.Tables(scal, tab, 256, 512)


def main():
    only = sys.argv[1:]
    want = lambda n: not only or n in only
    fid, w0wa = helpers.load_tables("fiducial").param(), helpers.load_tables("w0wa").param()
    grid512 = np.geomspace(1e-4, 10.0, 512)
    t = time.time()
    if want("config2_grid64_n265"):
        out = run_case(fid, D265, grid512[7::8], [1.0], 1e-4)
        np.savez_compressed(os.path.join(GOLD, "oracle_config2_grid64_n265.npz"), cosmology="fiducial", **out)
        print("config2_grid64", out["nsteps"].max(), "%.0fs" % (time.time() - t), flush=True)
    if want("config3_grid32_n265"):
        out = run_case(w0wa, D265, np.geomspace(1e-4, 10.0, 4096)[127::128], [1.0], 1e-4)
        np.savez_compressed(os.path.join(GOLD, "oracle_config3_grid32_n265.npz"), cosmology="w0wa", **out)
        print("config3_grid32", out["nsteps"].max(), "%.0fs" % (time.time() - t), flush=True)
    if want("config4_n265"):
        for i in range(3):
            tab = reference_tables(f"config4_{i}")
            out = run_case(tab.param(), D265, np.geomspace(1e-4, 10.0, 256)[15::16], [1.0], 1e-4)
            np.savez_compressed(os.path.join(GOLD, f"oracle_config4_{i}_n265.npz"), cosmology=f"config4_{i}", scalars=tab.scalars,
                                tables=tab.tables, nth=256, nnu=512, **out)
            print("config4", i, out["nsteps"].max(), "%.0fs" % (time.time() - t), flush=True)
    if want("config2_full512"):
        out = outputs_only(fid, D265, grid512, [1.0], 1e-4, max_steps=2048)
        np.savez_compressed(os.path.join(GOLD, "oracle_config2_full512.npz"), cosmology="fiducial", **out)
        print("config2_full512", int(out["nsteps"].sum()), "%.0fs" % (time.time() - t), flush=True)
    if want("class_grid512"):
        out = outputs_only(fid, D265, np.geomspace(1e-5, 10.0, 512), [0.01], 1e-4, max_steps=2048)
        np.savez_compressed(os.path.join(GOLD, "oracle_class_grid512.npz"), cosmology="fiducial", **out)
        print("class_grid512", int(out["nsteps"].sum()), "%.0fs" % (time.time() - t), flush=True)
    if want("converge_n265"):
        ks = grid512[15::16]
        res = {}
        for rt in (1e-4, 1e-5, 1e-7, 1e-8):
            o = outputs_only(fid, D265, ks, [1.0], rt, max_steps=32768, chunk=32)
            res[f"y_{rt:g}"] = o["y"]; res[f"nsteps_{rt:g}"] = o["nsteps"]
            print("converge", rt, int(o["nsteps"].max()), "%.0fs" % (time.time() - t), flush=True)
        np.savez_compressed(os.path.join(GOLD, "oracle_converge_n265.npz"), cosmology="fiducial", kmodes=ks, aexp_out=np.array([1.0]),
                            dims=np.array(D265), **res)


if __name__ == "__main__":
    main()
