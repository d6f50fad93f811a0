"""Writes a packed cosmology (tests/golden/tables_<name>.npz) as the flat binary examples/evolve_c_abi.c reads."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1] if len(sys.argv) > 1 else "fiducial"
out = sys.argv[2] if len(sys.argv) > 2 else "tables.bin"
z = np.load(os.path.join(ROOT, "tests", "golden", f"tables_{name}.npz"))
with open(out, "wb") as f:
    f.write(np.int32(z["nth"]).tobytes()); f.write(np.int32(z["nnu"]).tobytes())
    f.write(np.ascontiguousarray(z["scalars"], dtype=np.float64).tobytes())
    f.write(np.ascontiguousarray(z["tables"], dtype=np.float64).tobytes())
print("wrote", out)
