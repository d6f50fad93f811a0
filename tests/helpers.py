"""Test helpers: fixtures on disk -> packed tables / oracle-style ``param`` dicts."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle.background import Spline  # noqa: E402
from discoeb_b200._pack import SCALAR_KEYS, SPLINE_KEYS  # noqa: E402


class Tables:
    def __init__(self, scalars, tables, nth, nnu):
        self.scalars, self.tables, self.nth, self.nnu = scalars, tables, int(nth), int(nnu)

    def param(self):
        """Rebuild the dict the oracle consumes (splines carry the stored second derivatives)."""
        p = {k: float(self.scalars[i]) for i, k in enumerate(SCALAR_KEYS)}
        off = 0
        for j, key in enumerate(SPLINE_KEYS):
            n = self.nnu if j in (2, 3) else self.nth
            sp = Spline.__new__(Spline)
            sp.x = self.tables[off:off + n].copy()
            sp.y = self.tables[off + n:off + 2 * n].copy()
            sp.S = self.tables[off + 2 * n:off + 3 * n].copy()
            p[key] = sp
            off += 3 * n
        return p


def load_tables(name):
    z = np.load(os.path.join(GOLD, f"tables_{name}.npz"))
    return Tables(z["scalars"], z["tables"], z["nth"], z["nnu"])


def load_case(name):
    """``name`` -> tests/golden/oracle_<name>.npz (NumPy oracle); ``ref:<name>`` -> tests/golden/ref_<name>.npz, the same
    layout produced by the REFERENCE's own sources (tools/make_reference_fixtures.py)."""
    fn = f"ref_{name[4:]}.npz" if name.startswith("ref:") else f"oracle_{name}.npz"
    z = np.load(os.path.join(GOLD, fn), allow_pickle=False)
    return {k: z[k] for k in z.files}


def ref_cases():
    """Reference-produced fixtures present in tests/golden (``ref:<name>`` ids)."""
    import glob
    return tuple(sorted("ref:" + os.path.basename(f)[4:-4] for f in glob.glob(os.path.join(GOLD, "ref_*.npz"))))


# BASELINE.json configurations at their own size (tools/make_golden_baseline.py): replay fixtures with step traces
BASELINE_CASES = ("config2_grid64_n265", "config3_grid32_n265", "config4_0_n265", "config4_1_n265", "config4_2_n265")
CASES = ("default_n72", "config1_n111", "config2_n265", "w0wa_n72", "odd_dims_n43", "min_dims_n33", "many_out_n72")


def hierarchy_scales(b, dims):
    """Per-element magnitudes of a raw state ``b[..., n]`` for comparisons: a multipole is measured against the largest
    entry of ITS hierarchy (an oscillating moment passes through zero; its own value is no scale), fluid and metric
    variables against their own largest magnitude over the compared set."""
    lg, lp, lr, ln, nq = (int(v) for v in dims)
    ax = tuple(range(b.ndim - 1))
    sc = np.maximum(np.abs(b).max(axis=ax), 1e-300)
    ig, igp, ir, iq0 = 7, 8 + lg, 9 + lg + lp, 10 + lg + lp + lr
    for lo, hi in ((ig, igp), (igp, ir), (ir, iq0), (iq0, iq0 + nq * (ln + 1))):
        sc[lo:hi] = sc[lo:hi].max()
    sc[4] = max(sc[4], sc[6], sc[8])          # theta_c = 0 in synchronous gauge: round-off only
    return sc


def field_scaled_diff(a, b, dims=None):
    """max |a-b| per output field, relative to the largest magnitude that field takes over the
    compared set (fields such as theta_c are identically ~0, so a plain relative error is
    meaningless for them)."""
    if dims is not None and b.shape[-1] != 20:
        return np.abs(a - b).max(axis=tuple(range(b.ndim - 1))) / hierarchy_scales(b, dims)
    sc = np.maximum(np.abs(b).max(axis=tuple(range(b.ndim - 1))), 1e-300)
    # theta_c is identically zero in synchronous gauge (perturbations.py:269): it only carries
    # round-off, so it is measured against the baryon velocity that sits next to it
    # ... or the photon velocity (the baryon velocity itself is tiny for super-horizon modes)
    if b.shape[-1] == 20:
        sc[9] = max(sc[9], sc[11], sc[13])
        sc[11] = max(sc[11], sc[13])      # theta_b of a super-horizon mode is ~1e-8 of theta_gamma: same reasoning
    else:
        sc[4] = max(sc[4], sc[6], sc[8])
    return np.abs(a - b).max(axis=tuple(range(b.ndim - 1))) / sc
