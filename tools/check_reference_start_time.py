"""Runs the reference's own evolve_background + determine_starting_time (under tools/refshim) at num_thermo = 256 and 1024:
shows that with the default 256 thermo knots the start-time bisection fails for k >= 60/Mpc (BASELINE config "2a", kmax = 100).
TEST INFRASTRUCTURE; output committed as profiles/r2_reference_start_time_quirk.txt."""
import sys, os, warnings
import numpy as np
warnings.filterwarnings("ignore")
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "tools", "refshim")); sys.path.insert(0, "/root/reference/src"); sys.path.insert(0, ROOT)
import jax, jax.numpy as jnp
from discoeb.background import evolve_background
from discoeb import perturbations as RP
param = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09, n_s=0.96822, H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06)
for nt in (256, 1024):
    p = evolve_background(param=dict(param), thermo_module='RECFAST', num_thermo=nt)
    print("num_thermo", nt, "tau[0:3]", np.asarray(p['tau'])[:3], "aexp[0:3]", np.asarray(p['aexp'])[:3], "taumin", float(p['taumin']))
    for k in (1.0, 10.0, 40.0, 52.0, 60.0, 100.0):
        print("   k", k, "tau_start", float(RP.determine_starting_time(param=p, k=k)))
