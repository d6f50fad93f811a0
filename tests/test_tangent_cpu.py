"""Forward-tangent row on the CPU: the oracle's complex-step tangent is pinned against
frozen-step finite differences and the kernel SOURCE (tests/emu build of deb_core.cuh + deb_tangent.cuh) is
compared with it."""
import numpy as np
import pytest

import helpers
import parity_checks as pc
import oracle.discoeb_oracle as O
import oracle.discoeb_tangent as T
from oracle.background import Spline


def test_analytic_scale_factor_column_matches_complex_step():
    for name, dims in (("fiducial", (11, 11, 11, 8, 3)), ("w0wa", (5, 4, 6, 3, 4)), ("w0wa", (31, 31, 31, 31, 5))):
        p = helpers.load_tables(name).param()
        d = O.Dims(*dims)
        rng = np.random.default_rng(1)
        t = np.array([2.0, 30.0, 250.0, 900.0, 8000.0])
        k = np.geomspace(1e-3, 3, 5)
        y = rng.normal(size=(5, d.n))
        y[:, 0] = p["a_of_tau_spline"].evaluate(t)
        J = O.jacobian(t, y, p, k, d)                       # complex-step column
        da = O.rhs_da(t, y, p, k, d)                        # hand-derived column
        sc = np.abs(J[:, :, 0]).max(1, keepdims=True)
        assert (np.abs(da - J[:, :, 0]) / sc).max() < 1e-12


def _case_params(case):
    tab = helpers.Tables(case["scalars"], case["tables"], case["nth"], case["nnu"])
    p = tab.param()
    dps = []
    for d in range(case["d_scalars"].shape[0]):
        dt_ = helpers.Tables(case["d_scalars"][d], case["d_tables"][d], case["nth"], case["nnu"]).param()
        dps.append(dt_)
    return p, dps


def _shift(p, dp, eps):
    q = {}
    for k in T.SCALAR_KEYS:
        q[k] = float(p[k]) + eps * float(dp[k])
    for k in T.SPLINE_KEYS:
        s = Spline.__new__(Spline)
        s.x, s.y, s.S = p[k].x + eps * dp[k].x, p[k].y + eps * dp[k].y, p[k].S + eps * dp[k].S
        q[k] = s
    return q


def test_oracle_tangent_matches_frozen_step_finite_differences():
    """Pins the complex-step tangent: central differences of the (real) oracle along the SAME decisions."""
    case = pc.load_tangent_case("w0wa_n43")
    p, dps = _case_params(case)
    lg, lp, lr, ln, nq = (int(v) for v in case["dims"])
    kw = dict(aexp_out=case["aexp_out"], kmodes=case["kmodes"], rtol=float(case["rtol"]), atol=float(case["rtol"]),
              lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, max_steps=4096)
    out = T.evolve_perturbations_jvp(param=p, dparam=dps[0], **kw)
    np.testing.assert_allclose(out["dy"], case["dy"][0], rtol=1e-6, atol=1e-9 * np.abs(case["dy"][0]).max())
    rep = (out["rp_keep"], out["rp_fac"], out["nsteps"])
    eps = 1e-5
    op = T.evolve_perturbations_jvp(param=_shift(p, dps[0], eps), dparam={}, replay=rep, **kw)
    om = T.evolve_perturbations_jvp(param=_shift(p, dps[0], -eps), dparam={}, replay=rep, **kw)
    for f in ("tau_start", "y0", "y", "pk4"):
        fd = (op[f] - om[f]) / (2 * eps)
        ref = out["d" + f]
        if f == "y":
            for m in range(len(case["kmodes"])):
                assert pc.tangent_scaled_diff(fd[m], ref[m], out["y"][m]).max() < 1e-6
        else:
            assert np.abs(fd - ref).max() / np.abs(ref).max() < 1e-6, f


@pytest.mark.parametrize("name", pc.TANGENT_CASES)
def test_kernel_source_tangent_replay(emu_lib, name):
    worst = pc.check_tangent_replay(emu_lib, name)
    print(name, "worst scaled tangent deviation", worst)


@pytest.mark.parametrize("name", ("default_n72", "w0wa_n43", "kscaled_n72"))
def test_kernel_source_tangent_adaptive(emu_lib, name):
    pc.check_tangent_adaptive(emu_lib, name)


def test_kernel_source_tangent_properties(emu_lib):
    pc.check_tangent_properties(emu_lib)


def test_kernel_source_tangent_batch_of_cosmologies(emu_lib):
    pc.check_tangent_batch_of_cosmologies(emu_lib)
