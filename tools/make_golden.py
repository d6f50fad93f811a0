"""Generates the committed fixtures under tests/golden/ (run here, where the oracle is fast enough).

* ``tables_<name>.npz``  -- packed scalars/tables of a cosmology from the CPU table producer
  (oracle/background.py), so that GPU tests do not depend on re-integrating RECFAST.
* ``oracle_<case>.npz``  -- outputs of the NumPy oracle (oracle/discoeb_oracle.py) for small mode
  sets: 20-field outputs, raw states, attempted/accepted step counts and the full step trace
  (tnext, keep) that the replay parity test feeds back into the kernel.

The reference itself (JAX/diffrax) cannot be imported in this environment, so these are oracle
vectors, not reference vectors; the reference's own golden curves (CLASS_data.json,
RECFAST_DISCO_EB_data.json) are copied verbatim from /root/reference/tests/resources.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
import oracle.background as B  # noqa: E402
import oracle.discoeb_oracle as O  # noqa: E402
from discoeb_b200 import _pack  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

COSMOLOGIES = {
    "fiducial": {},                                                        # tests/test_perturbations.py:11-57
    "w0wa": dict(w_DE_0=-0.9, w_DE_a=0.1),                                 # BASELINE config 3
    "massless": dict(Nmnu=0, Neff=3.046, mnu=0.06),                        # BASELINE config 1
}

CASES = {
    # name: (cosmology, (lmaxg, lmaxgp, lmaxr, lmaxnu, nq), kmodes, aexp_out, rtol)
    "default_n72": ("fiducial", (11, 11, 11, 8, 3), np.geomspace(1e-4, 10.0, 8), [0.01, 0.5, 1.0], 1e-4),
    "config1_n111": ("massless", (16, 16, 16, 16, 3), np.geomspace(1e-4, 10.0, 6), [1.0], 1e-4),
    "config2_n265": ("fiducial", (31, 31, 31, 31, 5), np.array([1e-3, 0.03, 0.3, 3.0]), [0.01, 1.0], 1e-4),
    "w0wa_n72": ("w0wa", (11, 11, 11, 8, 3), np.array([1e-3, 0.1, 1.0]), [0.5, 1.0], 1e-5),
    "odd_dims_n43": ("fiducial", (5, 4, 6, 3, 4), np.array([2e-3, 0.2]), [0.1], 1e-3),
    "min_dims_n33": ("fiducial", (3, 3, 3, 3, 3), np.array([0.01, 0.3]), [0.05, 0.2, 1.0], 1e-4),       # tails of length 1
    "many_out_n72": ("w0wa", (11, 11, 11, 8, 3), np.array([1e-3, 0.05]), list(np.geomspace(1e-3, 1.0, 8)), 1e-5),
}


def helpers_param(name):
    """Rebuild the oracle param dict from the committed tables (keeps existing fixtures bit-stable)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    return helpers.load_tables(name).param()


def make_tables():
    params = {}
    for name, over in COSMOLOGIES.items():
        p = B.evolve_background(B.fiducial_param(**over))
        scal, tab, nth, nnu = _pack.pack_param(p)
        np.savez_compressed(os.path.join(GOLD, f"tables_{name}.npz"), scalars=scal, tables=tab, nth=nth, nnu=nnu)
        params[name] = p
    return params


def run_case(p, dims, ks, aout, rtol, max_steps=4096):
    d = O.Dims(*dims)
    aout = np.asarray(aout, dtype=np.float64)
    tau_out = p["tau_of_a_spline"].evaluate(aout)
    ts = 0.99 * np.minimum(tau_out.min(), O.determine_starting_time(p, ks))
    y0 = O.adiabatic_ics(ts, p, ks, d)
    tr = []
    ys, st, ns, na = O.integrate_modes(ts, tau_out.max(), y0, tau_out, p, ks, d, rtol, rtol, max_steps=max_steps, trace=tr)
    assert np.all(st == 0)
    M = len(ks)
    stride = int(ns.max())
    rp_t = np.zeros((M, stride))
    rp_k = np.zeros((M, stride), dtype=np.int32)
    rp_E = np.zeros((M, stride))
    cnt = np.zeros(M, dtype=np.int64)
    for act, tp, tn, E, keep in tr:
        for i, m in enumerate(act):
            rp_t[m, cnt[m]] = tn[i]
            rp_k[m, cnt[m]] = keep[i]
            rp_E[m, cnt[m]] = E[i]
            cnt[m] += 1
    assert np.all(cnt == ns)
    y20 = O.convert_to_output(ys, p, ks[:, None], d)
    return dict(kmodes=ks, aexp_out=aout, tau_out=tau_out, tau_start=ts, y0=y0, yfull=ys, y=y20, nsteps=ns.astype(np.int32),
                naccept=na.astype(np.int32), rp_tnext=rp_t, rp_keep=rp_k, rp_E=rp_E, dims=np.array(dims), rtol=rtol)


BATCHED = {
    # name: (cosmology, dims, (kmin, kmax, num_k), batch_size, aexp_out, rtol)
    "batched_n72": ("fiducial", (11, 11, 11, 8, 3), (1e-3, 0.3, 8), 4, [0.1, 1.0], 1e-4),
    "batched_lowk_n72": ("w0wa", (11, 11, 11, 8, 3), (1e-4, 3e-3, 8), 8, [0.5, 1.0], 1e-4),       # few steps: counts must match
    "batched_n265": ("fiducial", (31, 31, 31, 31, 5), (1e-3, 0.03, 16), 16, [1.0], 1e-4),          # 2 CTAs x 8 warps on the GPU
}


def run_batched_case(p, dims, kgrid, batch, aout, rtol):
    """Shared-step batches (oracle.evolve_perturbations_batched); the per-batch trace is replicated per mode so that
    the kernel's replay entry can follow it."""
    y, ks, info = O.evolve_perturbations_batched(param=p, aexp_out=aout, kmin=kgrid[0], kmax=kgrid[1], num_k=kgrid[2],
                                                 lmaxg=dims[0], lmaxgp=dims[1], lmaxr=dims[2], lmaxnu=dims[3], nqmax=dims[4],
                                                 rtol=rtol, atol=rtol, max_steps=4096, batch_size=batch, return_info=True)
    M = len(ks)
    stride = int(info["nsteps"].max())
    rp_t = np.zeros((M, stride)); rp_k = np.zeros((M, stride), dtype=np.int32)
    for b, tr in enumerate(info["traces"]):
        for s_, (tp, tn, E, keep) in enumerate(tr):
            rp_t[b * batch:(b + 1) * batch, s_] = tn
            rp_k[b * batch:(b + 1) * batch, s_] = keep
    aout = np.asarray(aout, dtype=np.float64)
    return dict(kmodes=ks, aexp_out=aout, tau_out=p["tau_of_a_spline"].evaluate(aout), y=y, yfull=info["yfull"],
                nsteps=np.repeat(info["nsteps"], batch).astype(np.int32), naccept=np.repeat(info["naccept"], batch).astype(np.int32),
                tau_start=np.repeat(info["tau_start"], batch), rp_tnext=rp_t, rp_keep=rp_k, dims=np.array(dims), rtol=rtol,
                batch_size=batch)


def main():
    only = sys.argv[1:]
    params = make_tables() if not only else {c: helpers_param(c) for c in COSMOLOGIES}
    for name, (cos, dims, ks, aout, rtol) in CASES.items():
        if only and name not in only:
            continue
        t = time.time()
        out = run_case(params[cos], dims, np.asarray(ks, dtype=np.float64), aout, rtol)
        np.savez_compressed(os.path.join(GOLD, f"oracle_{name}.npz"), cosmology=cos, **out)
        print(name, "steps", out["nsteps"], "%.1fs" % (time.time() - t), flush=True)
    for name, (cos, dims, kgrid, batch, aout, rtol) in BATCHED.items():
        if only and name not in only:
            continue
        t = time.time()
        out = run_batched_case(params[cos], dims, kgrid, batch, aout, rtol)
        np.savez_compressed(os.path.join(GOLD, f"oracle_{name}.npz"), cosmology=cos, **out)
        print(name, "steps", out["nsteps"][::batch], "%.1fs" % (time.time() - t), flush=True)


if __name__ == "__main__":
    main()
