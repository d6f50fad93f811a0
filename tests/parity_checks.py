"""Parity checks shared by the CPU suite (kernel source emulated on the CPU) and the GPU suite
(the CUDA library through its C-ABI).  Every check compares against the oracle
(oracle/discoeb_oracle.py) or its committed golden vectors (tests/golden/oracle_*.npz).

Tolerances (float64; north_star asks for 1e-5 relative on transfer functions / P(k)):

* prologue (start time, initial conditions): 1e-12 relative -- pure closed-form arithmetic.
* one Rodas5 step: 1e-7 of the state norm -- the kernel's structured solve (tri-diagonal tails +
  pivoted head) versus the oracle's dense LAPACK LU differ by cond(W)*eps.
* replay of the oracle's own step sequence over the whole integration: 1e-6 relative to each
  field's magnitude (1e-5 is the bar) -- the trajectory-level proof that RHS, Jacobian, solve,
  stage logic, output interpolation and conversion are the reference's.
* free-running adaptive solve: the step-size controller of the reference algorithm is chaotic
  under round-off once a mode takes more than ~100 steps (DESIGN.md "Parity"): two correct
  implementations diverge in their accept/reject sequence and then differ by O(10 rtol).  Hence
  modes of at most 100 steps whose attempted/accepted counts equal the oracle's must agree to 1e-5 (the
  error estimate itself carries ~1e-6 of round-off, which moves dt at the 1e-5 level); the others to
  50*rtol on the matter transfer functions, and at least half of all modes must be of the
  first kind.
"""
import os

import numpy as np

import helpers
import oracle.discoeb_oracle as O
from discoeb_b200 import _cabi

MATTER_FIELDS = (4, 6, 8, 10)      # delta_m, delta_bc, delta_c, delta_b  (what get_power is used on)


def dims_for(case, tab, nk, nout, **kw):
    lg, lp, lr, ln, nq = (int(v) for v in case["dims"])
    return _cabi.make_dims(ncosmo=1, nk=nk, nout=nout, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq,
                           nth=tab.nth, nnu=tab.nnu, max_steps=kw.pop("max_steps", 4096), **kw)


def case_tables(case, tables):
    """Packed tables of a case: the named committed set, or the set stored in the fixture itself (config-4 draws)."""
    if "scalars" in case and "tables" in case:
        return helpers.Tables(case["scalars"], case["tables"], int(case["nth"]), int(case["nnu"]))
    return tables[str(case["cosmology"])]


def check_prologue(lib, tables, name):
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    ks, aout = case["kmodes"], case["aexp_out"]
    dims = dims_for(case, tab, len(ks), len(aout))
    ts, y0 = lib.debug_ics(dims, tab.scalars, tab.tables, ks, aout)
    np.testing.assert_allclose(ts[0], case["tau_start"], rtol=1e-12)
    np.testing.assert_allclose(y0[0], case["y0"], rtol=1e-12, atol=0)


def check_single_step(lib, tables, name, seed=0):
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    p = tab.param()
    d = O.Dims(*(int(v) for v in case["dims"]))
    M = 6
    ks = np.geomspace(1e-3, 5.0, M)
    rng = np.random.default_rng(seed)
    t0 = np.array([2.0, 20.0, 120.0, 280.0, 900.0, 6000.0])
    t1 = t0 * (1 + np.array([0.2, 0.05, 0.02, 0.004, 0.03, 0.1]))
    y = rng.normal(size=(M, d.n))
    y[:, 0] = p["a_of_tau_spline"].evaluate(t0)
    dims = dims_for(case, tab, M, 1)
    y1, err = lib.debug_step(dims, tab.scalars, tab.tables, ks, t0, t1, y)
    y1o, erro = O.rodas5_step(t0, t1, y, p, ks, d)
    sc = np.abs(y1o).max(axis=1, keepdims=True)
    assert (np.abs(y1 - y1o) / sc).max() < 1e-7
    sce = np.abs(erro).max(axis=1, keepdims=True)
    assert (np.abs(err - erro) / sce).max() < 1e-5     # the error estimate is a small difference of large terms


def check_replay(lib, tables, name, tol=1e-6):
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    ks, aout = case["kmodes"], case["aexp_out"]
    ctrl = _cabi.make_ctrl(rtol=float(case["rtol"]), atol=float(case["rtol"]))
    for full in (False, True):
        dims = dims_for(case, tab, len(ks), len(aout), return_full=full)
        y, ns = lib.debug_replay(dims, ctrl, tab.scalars, tab.tables, ks, aout, case["rp_tnext"], case["rp_keep"], case["nsteps"])
        assert np.array_equal(ns[0], case["nsteps"])
        ref = case["yfull"] if full else case["y"]
        for m in range(len(ks)):          # per mode: the fields of different k differ by orders of magnitude
            assert helpers.field_scaled_diff(y[0, m], ref[m], dims=case["dims"]).max() < tol, (name, full, m)


def check_reference_step(lib, tables, name):
    """One Rodas5 step of the kernel against what the REFERENCE's own ``Rodas5Transformed.step`` returned for the same
    (t0, t1, y0) -- ref_<case>.npz: step_* (jacfwd Jacobian + LAPACK LU under tools/refshim).  Bars as check_single_step."""
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    ks = case["kmodes"]
    dims = dims_for(case, tab, len(ks), 1)
    y1, err = lib.debug_step(dims, tab.scalars, tab.tables, ks, case["step_t0"], case["step_t1"], case["rhs_state"])
    sc = np.abs(case["step_y1"]).max(axis=1, keepdims=True)
    assert (np.abs(y1 - case["step_y1"]) / sc).max() < 1e-7
    sce = np.abs(case["step_err"]).max(axis=1, keepdims=True)
    assert (np.abs(err - case["step_err"]) / sce).max() < 1e-5


def check_convergence(lib, tables, rtols=(1e-4, 1e-5), tight=1e-7):
    """VERDICT r1 item 1(b).  Truth = the oracle at rtol = atol = 1e-8 on 32 modes of BASELINE config 2 (n = 265, k up to
    10/Mpc; tests/golden/oracle_converge_n265.npz).  At every working tolerance the kernel's free-running error against
    the truth must not exceed 1.5 x the oracle's own (per field, worst mode; + 1e-7 floor) on the metric and matter
    fields 0-11 -- what the error norm controls (a, eta, delta_c, delta_b, theta_b, delta_gamma: perturbations.py:759)
    and what get_power reads -- and 3 x on the radiation / neutrino / dark-energy fields 12-19, which the norm does not
    see and whose z = 0 values at large k are tolerance-level noise in BOTH solvers (measured at rtol 1e-4: delta_m
    1.07e-4 kernel vs 1.08e-4 oracle; theta_gamma 3.8e-4 vs 2.2e-4; delta_gamma 0.49 vs 0.49 of its scale).  So the
    differences between kernel and oracle at rtol = 1e-4 are the SOLVER's tolerance-level error, present in the
    reference algorithm itself, not a defect of the structured solve.  At rtol = 1e-7 kernel and oracle agree to 1e-5 on
    the controlled fields (measured 7e-7 vs truth)."""
    import os
    z = np.load(os.path.join(helpers.GOLD, "oracle_converge_n265.npz"))
    case = {k: z[k] for k in z.files}
    tab = tables[str(case["cosmology"])]
    ks, aout = case["kmodes"], case["aexp_out"]
    truth = case["y_1e-08"]
    dims = dims_for(case, tab, len(ks), len(aout), max_steps=32768)
    report = {}
    for rt in tuple(rtols) + (tight,):
        ctrl = _cabi.make_ctrl(rtol=rt, atol=rt)
        out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout)
        assert np.all(out["status"] == 0)
        y, yo = out["y"][0], case[f"y_{rt:g}"]
        ek = np.array([helpers.field_scaled_diff(y[m], truth[m]) for m in range(len(ks))]).max(axis=0)
        eo = np.array([helpers.field_scaled_diff(yo[m], truth[m]) for m in range(len(ks))]).max(axis=0)
        eko = np.array([helpers.field_scaled_diff(y[m], yo[m]) for m in range(len(ks))]).max(axis=0)
        report[rt] = dict(delta_m_kernel_vs_truth=float(ek[4]), delta_m_oracle_vs_truth=float(eo[4]), kernel_vs_oracle_controlled=float(eko[:12].max()))
        if rt == tight:
            assert eko[:12].max() < 1e-5, (rt, eko)
        else:
            assert np.all(ek[:12] <= 1.5 * eo[:12] + 1e-7), (rt, ek, eo)
            assert np.all(ek[12:] <= 3.0 * eo[12:] + 1e-7), (rt, ek, eo)
    return report


def check_full_grid_parity(lib, tables, bar=1e-5):
    """All 512 modes of the bench workload (BASELINE config 2) against the oracle's committed P(k) and fields
    (tests/golden/oracle_config2_full512.npz): fraction of modes within `bar`, worst deviation, and the 50 rtol
    envelope every mode must satisfy (free-running solves: see check_adaptive)."""
    case = helpers.load_case("config2_full512")
    tab = tables["fiducial"]
    ks, aout, rtol = case["kmodes"], case["aexp_out"], float(case["rtol"])
    dims = dims_for(case, tab, len(ks), len(aout), max_steps=2048, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout, want_pk=True)
    assert np.all(out["status"] == 0)
    rel = np.abs(out["pk"][0, :, 0] / case["pk4"][:, 0] - 1)
    same = out["nsteps"][0] == case["nsteps"]
    assert rel.max() < 2 * 50 * rtol                     # P ~ delta^2
    assert rel[same & (case["nsteps"] <= 100)].max() < 2e-5
    return dict(frac_within_bar=float((rel < bar).mean()), max_rel=float(rel.max()), median_rel=float(np.median(rel)),
                same_step_counts=float(same.mean()))


def check_adaptive(lib, tables, name):
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    ks, aout, rtol = case["kmodes"], case["aexp_out"], float(case["rtol"])
    dims = dims_for(case, tab, len(ks), len(aout))
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout)
    assert np.all(out["status"] == 0)
    np.testing.assert_allclose(out["tau_out"][0], case["tau_out"], rtol=1e-14)
    same = out["nsteps"][0] == case["nsteps"]
    assert same.mean() >= 0.5, (out["nsteps"][0], case["nsteps"])
    y, ref = out["y"][0], case["y"]
    for m in range(len(ks)):
        rel = np.abs(y[m][:, MATTER_FIELDS] / ref[m][:, MATTER_FIELDS] - 1).max()
        if same[m] and np.array_equal(out["naccept"][0][m:m + 1], case["naccept"][m:m + 1]) and case["nsteps"][m] <= 100:
            assert helpers.field_scaled_diff(y[m], ref[m]).max() < 1e-5, (name, m)
        assert rel < 50 * rtol, (name, m, rel)
    return out


# ---------------------------------------------------------------------------------------------------
# forward tangents: kernel vs the complex-step oracle (tests/golden/tangent_*.npz,
# tools/make_golden_tangent.py).  Bars: replay of the oracle's step sequence 1e-6 (north_star: 1e-5) of each
# field's tangent scale; free-running, modes whose step counts equal the oracle's: 1e-5 for <= 100 steps.
# ---------------------------------------------------------------------------------------------------
TANGENT_CASES = ("default_n72", "w0wa_n43", "fisher_n265", "kscaled_n72", "config5_n265")


def load_tangent_case(name):
    import os
    z = np.load(os.path.join(helpers.GOLD, f"tangent_{name}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def tangent_dims(case, nk=None, **kw):
    lg, lp, lr, ln, nq = (int(v) for v in case["dims"])
    return _cabi.make_dims(ncosmo=1, nk=nk or len(case["kmodes"]), nout=len(case["aexp_out"]), lmaxg=lg, lmaxgp=lp, lmaxr=lr,
                           lmaxnu=ln, nqmax=nq, nth=int(case["nth"]), nnu=int(case["nnu"]), max_steps=kw.pop("max_steps", 4096),
                           ntan=kw.pop("ntan", case["d_scalars"].shape[0]), **kw)


def tangent_scaled_diff(dy, dref, ref):
    """max |dy - dref| per field relative to that field's tangent scale.  A field whose tangent is anomalously small
    (theta_c = 0, the scale factor under a post-processing parameter) is measured against |field| x the typical
    logarithmic derivative of the other fields instead."""
    ax = tuple(range(ref.ndim - 1))
    fmax = np.maximum(np.abs(ref).max(axis=ax), 1e-300)
    dmax = np.abs(dref).max(axis=ax)
    typical = np.median(dmax / fmax)
    sc = np.maximum(dmax, fmax * typical)
    nine = 9 if ref.shape[-1] == 20 else 4          # theta_c: round-off only (helpers.field_scaled_diff)
    sc[nine] = max(sc[nine], sc[nine + 2], sc[nine + 4])      # ... measured against the baryon / photon velocities
    return np.abs(dy - dref).max(axis=ax) / np.maximum(sc, 1e-300)


def check_tangent_replay(lib, name, tol=1e-6):
    case = load_tangent_case(name)
    ks, aout = case["kmodes"], case["aexp_out"]
    ctrl = _cabi.make_ctrl(rtol=float(case["rtol"]), atol=float(case["rtol"]))
    nt = case["d_scalars"].shape[0]
    worst = 0.0
    for full in (False, True):
        dims = tangent_dims(case, return_full=full)
        y, dy, dtau, ns = lib.debug_replay_tangent(dims, ctrl, case["scalars"][None], case["tables"][None], ks, aout,
                                                   case["d_scalars"][:, None], case["d_tables"][:, None], case["rp_tnext"],
                                                   case["rp_dtnext"], case["rp_keep"], case["nsteps"], d_kmodes=case.get("d_kmodes"))
        assert np.array_equal(ns[0], case["nsteps"])
        np.testing.assert_allclose(dtau[:, 0], case["dtau_out"], rtol=1e-9, atol=1e-9 * np.abs(case["dtau_out"]).max() + 1e-300)
        ref, dref = (case["yfull"], case["dyfull"]) if full else (case["y"], case["dy"])
        for m in range(len(ks)):
            assert helpers.field_scaled_diff(y[0, m], ref[m], dims=case["dims"]).max() < tol, (name, full, m)
            for d in range(nt):
                if not np.any(dref[d, m]):
                    assert not np.any(dy[d, 0, m]), (name, full, m, d)      # post-processing-only direction: exactly zero
                    continue
                e = tangent_scaled_diff(dy[d, 0, m], dref[d, m], ref[m]).max()
                worst = max(worst, e)
                assert e < tol, (name, full, m, d, e)
    return worst


# The tangent oracle pinned against the REFERENCE's functions (tools/make_reference_tangent.py): central differences with
# a Richardson step of the reference's ICs -> Rodas5Transformed.step along the recorded step sequence -> interpolation ->
# convert_to_output_variables -> get_power, run through tools/refshim.  Bar: 2e-6 of the field's tangent scale plus ten
# times the Richardson error estimate (differences of ~1e-4 relative steps resolve 1e-8 .. 1e-7 on well-scaled fields; a
# few tiny, cancellation-dominated fields are limited by the differencing, not by the tangent).
REFERENCE_TANGENT_CASES = ("default_n72", "w0wa_n43", "fisher_n265")


def load_reference_tangent(name):
    fn = os.path.join(helpers.GOLD, f"reference_tangent_{name}.npz")
    if not os.path.exists(fn):
        import pytest
        pytest.skip(f"{os.path.basename(fn)} not generated")
    z = np.load(fn, allow_pickle=False)
    return {k: z[k] for k in z.files}


def reference_tangent_diffs(name, dy, dyfull=None, dpk=None):
    """``dy[direction, mode, nout, 20]`` (fixture direction order) against the differenced reference; asserts the bar and
    returns the worst scaled deviation over the fields whose Richardson estimate is below 1e-8 of their scale."""
    case, ref = load_tangent_case(name), load_reference_tangent(name)
    worst = 0.0
    for im, m in enumerate(ref["modes"]):
        for idd, d in enumerate(ref["dir_index"]):
            for got, key, prim in ((dy, "dy", case["y"][m]), (dyfull, "dyfull", case["yfull"][m])):
                if got is None:
                    continue
                fd, err = ref[key][im, idd], ref[key + "_err"][im, idd]
                ax = tuple(range(prim.ndim - 1))
                # the scale of tangent_scaled_diff, kept per field
                fmax = np.maximum(np.abs(prim).max(axis=ax), 1e-300); dmax = np.abs(fd).max(axis=ax)
                sc = np.maximum(dmax, fmax * np.median(dmax / fmax))
                nine = 9 if prim.shape[-1] == 20 else 4
                sc[nine] = max(sc[nine], sc[nine + 2], sc[nine + 4])
                if prim.shape[-1] != 20:          # raw state: a multipole is measured against its hierarchy
                    sc = np.maximum(sc, helpers.hierarchy_scales(fd, case["dims"]))
                dev = np.abs(got[d, m] - fd)
                assert np.all(dev <= 2e-6 * sc + 10.0 * err), (name, key, int(m), int(d), float((dev / sc).max()))
                ok = err.max(axis=ax) < 1e-8 * sc           # fields the differencing resolves well
                if np.any(ok):
                    worst = max(worst, float((dev.max(axis=ax) / sc)[ok].max()))
            if dpk is not None:
                rel = np.abs(dpk[d, m] / ref["dpk4"][im, idd] - 1.0)
                assert np.all(rel <= 2e-6 + 10.0 * np.abs(ref["dpk4_err"][im, idd] / ref["dpk4"][im, idd])), (name, "dpk", int(m), int(d), rel)
    return worst


def check_tangent_replay_vs_reference(lib, name):
    """The kernel's tangent replay against the differenced reference directly."""
    case = load_tangent_case(name)
    ctrl = _cabi.make_ctrl(rtol=float(case["rtol"]), atol=float(case["rtol"]))
    got = {}
    for full in (False, True):
        dims = tangent_dims(case, return_full=full)
        _, dy, _, _ = lib.debug_replay_tangent(dims, ctrl, case["scalars"][None], case["tables"][None], case["kmodes"], case["aexp_out"],
                                               case["d_scalars"][:, None], case["d_tables"][:, None], case["rp_tnext"],
                                               case["rp_dtnext"], case["rp_keep"], case["nsteps"], d_kmodes=case.get("d_kmodes"))
        got[full] = dy[:, 0]
    return reference_tangent_diffs(name, got[False], got[True])


def check_tangent_adaptive(lib, name):
    """Free-running: the public tangent entry (controller on) against the oracle, incl. P(k) and its tangent."""
    case = load_tangent_case(name)
    ks, aout, rtol = case["kmodes"], case["aexp_out"], float(case["rtol"])
    nt = case["d_scalars"].shape[0]
    dims = tangent_dims(case, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    out = lib.evolve_tangent_host(dims, ctrl, case["scalars"][None], case["tables"][None], ks, aout, case["d_scalars"][:, None],
                                  case["d_tables"][:, None], want_pk=True, d_kmodes=case.get("d_kmodes"))
    assert np.all(out["status"] == 0)
    np.testing.assert_allclose(out["tau_out"][0], case["tau_out"], rtol=1e-13)
    # the primal half must be what the primal entry returns for the same inputs (same kernel code path modulo the
    # helper-warp variant: 50 rtol; bitwise when that variant is not selected is checked on the GPU suite)
    same = (out["nsteps"][0] == case["nsteps"]) & (out["naccept"][0] == case["naccept"])
    assert same[case["nsteps"] <= 100].all(), (out["nsteps"][0], case["nsteps"])      # short modes: identical step counts
    for m in range(len(ks)):
        relp = np.abs(out["pk"][0, m] / case["pk4"][m] - 1).max()
        assert relp < 400 * rtol, (name, m, relp)      # P ~ delta^2: twice the 50..200 rtol global error of a free-running solve
        for d in range(nt):
            dref = case["dpk4"][d, m]
            rel = np.abs(out["dpk"][d, 0, m] - dref).max() / (np.abs(dref).max() + 1e-9 * np.abs(case["pk4"][m]).max())
            tight = same[m] and case["nsteps"][m] <= 100
            assert rel < (1e-5 if tight else 2000 * rtol), (name, m, d, rel)
            if tight and np.any(case["dy"][d, m]):
                assert tangent_scaled_diff(out["dy"][d, 0, m], case["dy"][d, m], case["y"][m]).max() < 1e-5, (name, m, d)
    return out


def check_tangent_properties(lib, name="default_n72", nk=16):
    """Size-independent properties of the tangent map: linearity in the seed, d P/d A_s = P/A_s,
    d P/d n_s = P log(k/k_p), zero seed -> zero tangent, and the primal half equals the plain solve."""
    case = load_tangent_case(name)
    aout, rtol = case["aexp_out"], float(case["rtol"])
    ks = np.geomspace(1e-3, 1.0, nk)
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    sc, tb = case["scalars"], case["tables"]
    d1s, d1t, d2s, d2t = case["d_scalars"][0], case["d_tables"][0], case["d_scalars"][1], case["d_tables"][1]
    dAs = np.zeros_like(sc); dAs[16] = 1.0
    dns = np.zeros_like(sc); dns[17] = 1.0
    zt = np.zeros_like(tb)
    seeds_s = np.stack([d1s, d2s, 0.3 * d1s - 1.7 * d2s, dAs, dns, np.zeros_like(sc)])
    seeds_t = np.stack([d1t, d2t, 0.3 * d1t - 1.7 * d2t, zt, zt, zt])
    dims = tangent_dims(case, nk=nk, ntan=len(seeds_s), power_idx=4)
    out = lib.evolve_tangent_host(dims, ctrl, sc[None], tb[None], ks, aout, seeds_s[:, None], seeds_t[:, None], want_pk=True)
    assert np.all(out["status"] == 0)
    dy, dpk, pk = out["dy"][:, 0], out["dpk"][:, 0], out["pk"][0]
    lin = 0.3 * dy[0] - 1.7 * dy[1]
    scale = np.abs(dy[0]).max(axis=(0, 1)) + np.abs(dy[1]).max(axis=(0, 1)) + 1e-300
    assert (np.abs(dy[2] - lin).max(axis=(0, 1)) / scale).max() < 1e-9
    np.testing.assert_allclose(dpk[3], pk / sc[16], rtol=1e-12)
    np.testing.assert_allclose(dpk[4], pk * np.log(ks / sc[18])[:, None], rtol=1e-10, atol=1e-12 * np.abs(pk).max())
    assert not np.any(dy[3]) and not np.any(dy[4]) and not np.any(dy[5]) and not np.any(dpk[5])
    return out


def check_tangent_batch_of_cosmologies(lib, nk=6):
    """[ntan, ncosmo, ...] indexing: two cosmologies x two directions in one launch equal the four single launches;
    per-cosmology k grids and the raw-state output included."""
    a, b = load_tangent_case("default_n72"), load_tangent_case("w0wa_n43")
    # same table sizes, different cosmologies (fiducial / w0wa); use the n=72 layout for both
    sc = np.stack([a["scalars"], b["scalars"]]); tb = np.stack([a["tables"], b["tables"]])
    ds = np.stack([np.stack([a["d_scalars"][0], b["d_scalars"][0]]), np.stack([a["d_scalars"][1], b["d_scalars"][1]])])
    dt = np.stack([np.stack([a["d_tables"][0], b["d_tables"][0]]), np.stack([a["d_tables"][1], b["d_tables"][1]])])
    ks = np.stack([np.geomspace(1e-3, 0.3, nk), np.geomspace(2e-3, 0.5, nk)])
    aout = np.array([0.2, 1.0])
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    for full in (False, True):
        dims = tangent_dims(a, nk=nk, ntan=2, power_idx=-1 if full else 6, k_per_cosmo=True, return_full=full)
        dims.ncosmo = 2
        out = lib.evolve_tangent_host(dims, ctrl, sc, tb, ks, aout, ds, dt, want_pk=not full)
        assert np.all(out["status"] == 0)
        for c in range(2):
            for d in range(2):
                d1 = tangent_dims(a, nk=nk, ntan=1, power_idx=-1 if full else 6, return_full=full)
                one = lib.evolve_tangent_host(d1, ctrl, sc[c:c + 1], tb[c:c + 1], ks[c], aout, ds[d:d + 1, c:c + 1], dt[d:d + 1, c:c + 1],
                                              want_pk=not full)
                assert np.array_equal(one["nsteps"][0], out["nsteps"][c])
                np.testing.assert_array_equal(one["y"][0], out["y"][c])
                np.testing.assert_array_equal(one["dy"][0, 0], out["dy"][d, c])
                np.testing.assert_array_equal(one["dtau_out"][0, 0], out["dtau_out"][d, c])
                if not full:
                    np.testing.assert_array_equal(one["dpk"][0, 0], out["dpk"][d, c])


def check_edge_shapes(lib, tables):
    """Shapes the golden cases do not cover: the largest supported hierarchy (n = 337, 11 elements per lane), four
    momentum bins, a single mode with a single output, and a non-finite input (status 2, as diffrax's non-finite result)."""
    tab = tables["fiducial"]
    p = tab.param()
    for dims5 in ((40, 40, 40, 40, 5), (12, 9, 10, 7, 4)):
        d = O.Dims(*dims5)
        M = 4
        ks = np.geomspace(2e-3, 2.0, M)
        rng = np.random.default_rng(3)
        t0 = np.array([3.0, 40.0, 300.0, 5000.0])
        t1 = t0 * 1.03
        y = rng.normal(size=(M, d.n))
        y[:, 0] = p["a_of_tau_spline"].evaluate(t0)
        dims = _cabi.make_dims(ncosmo=1, nk=M, nout=1, lmaxg=dims5[0], lmaxgp=dims5[1], lmaxr=dims5[2], lmaxnu=dims5[3],
                               nqmax=dims5[4], nth=tab.nth, nnu=tab.nnu, max_steps=4096)
        y1, err = lib.debug_step(dims, tab.scalars, tab.tables, ks, t0, t1, y)
        y1o, erro = O.rodas5_step(t0, t1, y, p, ks, d)
        assert (np.abs(y1 - y1o) / np.abs(y1o).max(axis=1, keepdims=True)).max() < 1e-7, dims5
    # one mode, one output, largest hierarchy, free-running
    dims = _cabi.make_dims(ncosmo=1, nk=1, nout=1, lmaxg=40, lmaxgp=40, lmaxr=40, lmaxnu=40, nqmax=5, nth=tab.nth, nnu=tab.nnu,
                           max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    out = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], np.array([0.02]), np.array([1.0]), want_pk=True)
    assert out["status"][0, 0] == 0 and out["y"].shape == (1, 1, 1, 20) and out["pk"][0, 0, 0] > 0
    dims31 = _cabi.make_dims(ncosmo=1, nk=1, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu,
                             max_steps=4096, power_idx=4)
    out31 = lib.evolve_host(dims31, ctrl, tab.scalars[None], tab.tables[None], np.array([0.02]), np.array([1.0]), want_pk=True)
    assert abs(out["pk"][0, 0, 0] / out31["pk"][0, 0, 0] - 1) < 1e-3          # the cut-off does not matter at k = 0.02
    # step budget exhausted (diffrax: max_steps reached with throw=False): status 1, exactly max_steps attempts, on every variant's
    # layout (2 modes: team; the short mode still finishes)
    dims = _cabi.make_dims(ncosmo=1, nk=2, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth, nnu=tab.nnu,
                           max_steps=25)
    outm = lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], np.array([1e-4, 1.0]), np.array([1.0]))
    assert outm["status"][0].tolist() == [0, 1] and outm["nsteps"][0, 1] == 25 and outm["nsteps"][0, 0] < 25
    # non-finite input
    dims = _cabi.make_dims(ncosmo=1, nk=2, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth, nnu=tab.nnu,
                           max_steps=256)
    for idx in (4, 15):                                                      # grhom (every coefficient), taumin (the start time)
        bad = tab.scalars.copy()
        bad[idx] = np.nan
        outb = lib.evolve_host(dims, ctrl, bad[None], tab.tables[None], np.array([0.01, 0.1]), np.array([1.0]))
        assert np.all(outb["status"] != 0), idx


# ---------------------------------------------------------------------------------------------------
# batched variant: one shared adaptive step per batch of modes
# (Rodas5Batched, evolve_modes_batched).  Oracle: oracle.evolve_perturbations_batched -> tests/golden/oracle_batched_*.npz
# ---------------------------------------------------------------------------------------------------
BATCHED_CASES = ("batched_n72", "batched_lowk_n72", "batched_n265")


def check_batched_replay(lib, tables, name, tol=1e-6):
    """Shared start time + the oracle's shared step sequence: whole trajectories at round-off level."""
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    ks, aout, B = case["kmodes"], case["aexp_out"], int(case["batch_size"])
    ctrl = _cabi.make_ctrl(rtol=float(case["rtol"]), atol=float(case["rtol"]))
    for full in (False, True):
        dims = dims_for(case, tab, len(ks), len(aout), return_full=full, batch_size=B)
        y, ns = lib.debug_replay(dims, ctrl, tab.scalars, tab.tables, ks, aout, case["rp_tnext"], case["rp_keep"], case["nsteps"])
        assert np.array_equal(ns[0], case["nsteps"])
        ref = case["yfull"] if full else case["y"]
        for m in range(len(ks)):
            assert helpers.field_scaled_diff(y[0, m], ref[m], dims=case["dims"]).max() < tol, (name, full, m)


def check_batched_adaptive(lib, tables, name):
    """Free-running: every mode of a batch reports the same step counts; batches of at most 100 steps must reproduce
    the oracle's counts and agree to 1e-5, longer ones to 50 rtol on the matter fields (DESIGN.md "Parity")."""
    case = helpers.load_case(name)
    tab = case_tables(case, tables)
    ks, aout, B, rtol = case["kmodes"], case["aexp_out"], int(case["batch_size"]), float(case["rtol"])
    dims = dims_for(case, tab, len(ks), len(aout), batch_size=B)
    out = lib.evolve_host(dims, _cabi.make_ctrl(rtol=rtol, atol=rtol), tab.scalars[None], tab.tables[None], ks, aout)
    assert np.all(out["status"] == 0)
    ns, na = out["nsteps"][0].reshape(-1, B), out["naccept"][0].reshape(-1, B)
    assert np.all(ns == ns[:, :1]) and np.all(na == na[:, :1])                 # lock-step inside a batch
    for b in range(ns.shape[0]):
        sl = slice(b * B, (b + 1) * B)
        short = case["nsteps"][b * B] <= 100
        if short:
            assert ns[b, 0] == case["nsteps"][b * B] and na[b, 0] == case["naccept"][b * B], (name, b, ns[b, 0], case["nsteps"][b * B])
        for m in range(sl.start, sl.stop):
            if short:
                assert helpers.field_scaled_diff(out["y"][0, m], case["y"][m]).max() < 1e-5, (name, m)
            rel = np.abs(out["y"][0, m][:, MATTER_FIELDS] / case["y"][m][:, MATTER_FIELDS] - 1).max()
            assert rel < 50 * rtol, (name, m, rel)
    return out
