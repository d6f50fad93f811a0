"""Cross-check of the oracle (and, on a GPU box, the CUDA path) against the REAL reference.

Only runs where `jax`, `diffrax`, `equinox`, `jax_cosmo` and the reference package `discoeb` are
importable -- none are installable in the build environment (no wheels, no network), which is
why DESIGN.md says "1e-5 parity unpinned".  Usage:

    PYTHONPATH=/path/to/DISCO-EB/src python tools/crosscheck_jax.py

It feeds the reference's own `evolve_background` output to (i) the reference solver, (ii) the NumPy
oracle and (iii) the CUDA library if a GPU is visible, on a handful of modes, and reports the
attempted-step counts and the maximum relative deviation of the matter transfer functions, plus
wall times of the reference's CPU path (the CPU baseline bench.py cannot take here).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))

try:
    import jax
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    from discoeb.background import evolve_background
    from discoeb.perturbations import evolve_perturbations as ref_evolve
except Exception as exc:  # pragma: no cover - environment dependent
    print(f"crosscheck_jax: the reference is not importable here ({type(exc).__name__}: {exc}); nothing to do.")
    sys.exit(0)

import oracle.background as B  # noqa: E402
import oracle.discoeb_oracle as O  # noqa: E402
from oracle.background import Spline  # noqa: E402


def to_oracle_param(p):
    """Reference param dict (JAX pytrees) -> oracle dict with NumPy splines carrying the SAME S."""
    q = {k: (float(v) if np.ndim(v) == 0 else np.asarray(v)) for k, v in p.items() if not hasattr(v, "_x_")}
    for k, v in p.items():
        if hasattr(v, "_x_"):
            s = Spline.__new__(Spline)
            s.x, s.y, s.S = np.asarray(v._x_), np.asarray(v._y_), np.asarray(v._S_full_)
            q[k] = s
    return q


def main():
    param = {k: v for k, v in B.fiducial_param().items()}
    param = evolve_background(param=param, thermo_module="RECFAST")
    cfg = dict(aexp_out=jnp.array([1.0]), kmin=1e-3, kmax=1.0, num_k=8, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3,
               rtol=1e-5, atol=1e-5, max_steps=8192)
    t = time.time()
    out = ref_evolve(param=param, **cfg)
    y_ref, k_ref = np.asarray(out[0]), np.asarray(out[1])
    print(f"reference: {time.time() - t:.1f} s (includes tracing)")
    po = to_oracle_param(param)
    ocfg = dict(cfg)
    ocfg["aexp_out"] = [1.0]
    y_or, k_or, _, info = O.evolve_perturbations(param=po, return_info=True, **ocfg)
    mf = [4, 6, 8, 10]
    print("oracle attempted steps", info["nsteps"])
    print("oracle vs reference, matter fields, max rel:", np.abs(y_or[..., mf] / y_ref[..., mf] - 1).max())
    try:
        from discoeb_b200.perturbations import evolve_perturbations as gpu_evolve
        y_gpu, _, _, ginfo = gpu_evolve(param=po, return_info=True, **ocfg)
        print("cuda attempted steps  ", ginfo["nsteps"])
        print("cuda vs reference, matter fields, max rel:", np.abs(y_gpu[..., mf] / y_ref[..., mf] - 1).max())
    except Exception as exc:  # pragma: no cover
        print("cuda path skipped:", exc)


if __name__ == "__main__":
    main()
