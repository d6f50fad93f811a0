"""One tangent launch at n=265 (target of ncu captures): argv = num_k, ntan."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
from discoeb_b200 import _cabi
lib = _cabi.default_library()
z = np.load(os.path.join(ROOT, "tests", "golden", "fisher_seeds.npz"))
nk, ntan = int(sys.argv[1]), int(sys.argv[2])
ks = np.geomspace(1e-4, 10.0, nk)
dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=2, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=int(z["nth"]), nnu=int(z["nnu"]),
                       max_steps=4096, power_idx=4, ntan=ntan)
ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
out = lib.evolve_tangent_host(dims, ctrl, z["scalars"][None], z["tables"][None], ks, np.array([0.5, 1.0]), z["d_scalars"][:ntan, None],
                              z["d_tables"][:ntan, None], want_pk=True)
print("kernel_ms", out["kernel_ms"], "steps", int(out["nsteps"].sum()), "status", np.unique(out["status"]))
