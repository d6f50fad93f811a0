"""tests/golden/fisher_seeds.npz: packed tables of the Fisher-notebook fiducial (nb_discoeb_fisher_forecast.ipynb
cell 3: Omegam 0.32, Omegab 0.05, h 0.67, n_s 0.96, w0 -0.9999, wa 0, cs2 0.9999) and the tangent seeds of its 7
forecast parameters (Omegam, Omegab, w0, wa, h, n_s, A_s) from central differences of the CPU table producer --
the inputs of BASELINE config 5."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_tangent import param_tangent, pack_tangent  # noqa: E402
from discoeb_b200 import _pack  # noqa: E402

FID = dict(Omegam=0.32, Omegab=0.05, H0=67.0, n_s=0.96, w_DE_0=-0.9999, w_DE_a=0.0, cs2_DE=0.9999)
KEYS = ("Omegam", "Omegab", "w_DE_0", "w_DE_a", "H0", "n_s", "A_s")

if __name__ == "__main__":
    ds, dt = [], []
    for key in KEYS:
        p, dp = param_tangent(key, **FID)
        if key == "H0":                 # the forecast parameter is h
            dp = {k: (100.0 * v if not hasattr(v, "x") else v) for k, v in dp.items()}
            for k in _pack.SPLINE_KEYS:
                dp[k].x, dp[k].y, dp[k].S = 100.0 * dp[k].x, 100.0 * dp[k].y, 100.0 * dp[k].S
        a, b = pack_tangent(dp)
        ds.append(a)
        dt.append(b)
    scal, tab, nth, nnu = _pack.pack_param(p)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fisher_seeds.npz"), scalars=scal, tables=tab, nth=nth, nnu=nnu,
                        d_scalars=np.stack(ds), d_tables=np.stack(dt), keys=np.array(KEYS))
    print("ok", np.stack(ds).shape, np.stack(dt).shape)
