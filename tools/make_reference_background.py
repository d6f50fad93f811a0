"""Runs the reference's own ``evolve_background(param, thermo_module='RECFAST')`` (/root/reference/src/discoeb) under
tools/refshim and commits the tables evolve_perturbations reads -> tests/golden/reference_background.npz (TEST
INFRASTRUCTURE; see tools/refshim/README.md for what is executed and what is restated).

Cosmologies: the reference's test fiducial, BASELINE config 3 (w0wa), config 1 (massless), the Fisher notebook
fiducial, and the first BASELINE config-4 draws of numpy.random.default_rng(0)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
warnings.filterwarnings("ignore")

import jax  # noqa: E402 (refshim)
from discoeb.background import evolve_background  # noqa: E402 (the reference)
from discoeb_b200.background import BG_KEYS, config4_draws, pack_background_input  # noqa: E402

SPLINES = ("cs2a_of_loga_spline", "xe_of_loga_spline", "logrhonu_of_loga_spline", "logpnu_of_loga_spline", "a_of_tau_spline",
           "xe_of_tau_spline", "tau_of_a_spline")
SCALARS = ("Omegam", "Omegab", "OmegaDE", "Omegak", "grhom", "grhog", "grhor", "Neff", "Nmnu", "amnu", "w_DE_0", "w_DE_a", "cs2_DE",
           "YHe", "H0", "taumin")


def fiducial(**over):
    p = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09, n_s=0.96822,
             H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06, k_p=0.05)
    p.update(over)
    return p


# everything else evolve_background leaves in param (background.py:238-342): species fractions, sound speed, matter
# temperature, the tau-splines of a c_s^2 and a T_m, the pseudo-pressure table, optical depth and visibility
EXTRA_ARRAYS = ("xeHI", "xeHeI", "xeHeII", "cs2", "Tm", "xeprime_recf", "xeprime", "opac", "optical_depth", "gvis", "gvisprime", "gvispprime")
EXTRA_SPLINES = ("cs2a_of_tau_spline", "tempba_of_tau_spline", "logppseudonu_of_loga_spline")


def main():
    cases = {"fiducial": fiducial(), "w0wa": fiducial(w_DE_0=-0.9, w_DE_a=0.1), "massless": fiducial(Nmnu=0, Neff=3.046),
             "fisher": fiducial(Omegam=0.32, Omegab=0.05, H0=67.0, n_s=0.96, w_DE_0=-0.9999, cs2_DE=0.9999)}
    for i, d in enumerate(config4_draws(3)):
        cases[f"config4_{i}"] = fiducial(**d)
    out = {"names": np.array(list(cases))}
    ext = {"names": np.array(list(cases))}
    for name, p in cases.items():
        inp = pack_background_input(p)
        q = evolve_background(param=dict(p), thermo_module="RECFAST")
        out[f"{name}_in"] = inp
        out[f"{name}_scalars"] = np.array([float(q[k]) for k in SCALARS])
        for sp in SPLINES:
            out[f"{name}_{sp}_x"] = np.asarray(q[sp]._x_)
            out[f"{name}_{sp}_y"] = np.asarray(q[sp]._y_)
            out[f"{name}_{sp}_S"] = np.asarray(q[sp]._S_full_)
        for key in EXTRA_ARRAYS:
            ext[f"{name}_{key}"] = np.asarray(q[key], dtype=np.float64)
        for sp in EXTRA_SPLINES:
            ext[f"{name}_{sp}_y"] = np.asarray(q[sp]._y_)
            ext[f"{name}_{sp}_S"] = np.asarray(q[sp]._S_full_)
        ext[f"{name}_taumax"] = float(q["taumax"]); ext[f"{name}_adotrad"] = float(q["adotrad"]); ext[f"{name}_Omegamnu"] = float(q["Omegamnu"])
        print(name, "OmegaDE", float(q["OmegaDE"]), "xe[120]", float(np.asarray(q["xe"])[120]), flush=True)
    if "--extras-only" not in sys.argv:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_background.npz"), **out)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_background_extras.npz"), **ext)


if __name__ == "__main__":
    main()
