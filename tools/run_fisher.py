"""BASELINE config 5 on one GPU: P(k) and its Jacobian w.r.t. 7 cosmological parameters (forward tangents),
num_k=512, a_out = [0.5, 1.0], rtol=atol=1e-4, for the default hierarchy (n=72, as in the Fisher notebook) and
for n=265.  Prints one JSON line per shape."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
from discoeb_b200 import _cabi
lib = _cabi.default_library()
z = np.load(os.path.join(ROOT, "tests", "golden", "fisher_seeds.npz"))
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ks = np.geomspace(1e-4, 10.0, nk)
aout = np.array([0.5, 1.0])
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
for dims5 in ((11, 11, 11, 8, 3), (31, 31, 31, 31, 5)):
    lg, lp, lr, ln, nq = dims5
    n = lib.nvar(*dims5)
    for ntan in (0, 1, 7):
        dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=2, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=int(z["nth"]),
                               nnu=int(z["nnu"]), max_steps=4096, power_idx=4, ntan=ntan)
        best, wall = 1e30, 1e30
        for rep in range(3):
            t = time.perf_counter()
            if ntan == 0:
                out = lib.evolve_host(dims, ctrl, z["scalars"][None], z["tables"][None], ks, aout, want_pk=True)
            else:
                out = lib.evolve_tangent_host(dims, ctrl, z["scalars"][None], z["tables"][None], ks, aout,
                                              z["d_scalars"][:ntan, None], z["d_tables"][:ntan, None], want_pk=True)
            wall = min(wall, 1e3 * (time.perf_counter() - t))
            best = min(best, out["kernel_ms"])
        ok = bool(np.all(out["status"] == 0))
        line = dict(workload=f"config5: n={n}, num_k={nk}, a_out=[0.5,1], ntan={ntan}", kernel_ms=best, e2e_ms=wall, ok=ok,
                    steps=int(out["nsteps"].sum()), modes_per_s=nk / (best * 1e-3))
        if ntan:
            dlnP = out["dpk"][:, 0, :, -1] / out["pk"][0, :, -1]
            line["dlnP_dtheta_at_kmid"] = [float(v) for v in dlnP[:, nk // 2]]
            line["finite"] = bool(np.all(np.isfinite(out["dy"])))
        print(json.dumps(line), flush=True)
