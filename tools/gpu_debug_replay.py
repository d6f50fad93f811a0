import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers, parity_checks as pc
from discoeb_b200 import _cabi
lib = _cabi.default_library()
np.set_printoptions(linewidth=200, precision=3)
tables = {n: helpers.load_tables(n) for n in ("fiducial", "w0wa", "massless")}
for name in helpers.CASES:
    case = helpers.load_case(name); tab = tables[str(case["cosmology"])]
    ks, aout = case["kmodes"], case["aexp_out"]
    ctrl = _cabi.make_ctrl(rtol=float(case["rtol"]), atol=float(case["rtol"]))
    for full in (False, True):
        dims = pc.dims_for(case, tab, len(ks), len(aout), return_full=full)
        y, ns = lib.debug_replay(dims, ctrl, tab.scalars, tab.tables, ks, aout, case["rp_tnext"], case["rp_keep"], case["nsteps"])
        ref = case["yfull"] if full else case["y"]
        for m in range(len(ks)):
            d = helpers.field_scaled_diff(y[0, m], ref[m])
            flag = "FAIL" if d.max() > 1e-6 else ""
            print(name, "full" if full else "y20", "mode", m, "k", ks[m], "max", d.max(), "argmax", d.argmax(), flag)
            if flag:
                print("   diffs", d)
                print("   y  ", y[0, m][:, d.argmax()], "ref", ref[m][:, d.argmax()])
