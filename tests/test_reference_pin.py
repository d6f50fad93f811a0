"""CPU suite, part 0: the ORACLE and the KERNEL SOURCE against the REFERENCE ITSELF.

tests/golden/ref_<case>.npz hold what the unmodified reference sources (/root/reference/src/discoeb) returned when run
in the builder container under the NumPy stand-ins of tools/refshim (tools/make_reference_fixtures.py): every function
of the hot path executes the reference's own Python -- ``model_synchronous``, ``Rodas5Transformed.step`` (jacfwd as an
exact complex-step Jacobian, LAPACK LU), ``determine_starting_time``, ``adiabatic_ics_one_mode``,
``convert_to_output_variables``, ``get_power``, the ``spline_interpolation`` constructor -- and only diffrax's adaptive
loop is a restatement.  These tests pin

* the oracle (oracle/discoeb_oracle.py) function by function at round-off level, and over whole solves,
* the kernel source (CPU build, tests/emu; the CUDA build runs the same checks in tests/test_gpu_parity.py) by
  replaying it along the REFERENCE's step sequence and comparing with the REFERENCE's outputs (1e-6).

Nothing here reads /root/reference.
"""
import numpy as np
import pytest

import helpers
import parity_checks as pc
import oracle.discoeb_oracle as O

REF = helpers.ref_cases()
assert REF, "tests/golden/ref_*.npz missing (tools/make_reference_fixtures.py)"


def _setup(name, tables):
    case = helpers.load_case(name)
    tab = tables[str(case["cosmology"])]
    return case, tab.param(), O.Dims(*(int(v) for v in case["dims"]))


@pytest.mark.parametrize("name", REF)
def test_spline_constructor_is_the_references(name):
    """The reference's spline constructor, fed the committed knots, reproduced the stored second derivatives."""
    assert float(helpers.load_case(name)["spline_S_maxdiff"]) < 1e-13


@pytest.mark.parametrize("name", REF)
def test_oracle_prologue_vs_reference(name, tables):
    case, p, d = _setup(name, tables)
    ks = case["kmodes"]
    ts = 0.99 * np.minimum(case["tau_out"].min(), O.determine_starting_time(p, ks))
    np.testing.assert_allclose(ts, case["tau_start"], rtol=1e-14)
    np.testing.assert_allclose(O.adiabatic_ics(case["tau_start"], p, ks, d), case["y0"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("name", REF)
def test_oracle_rhs_vs_reference(name, tables):
    case, p, d = _setup(name, tables)
    f = O.rhs(case["rhs_tau"], case["rhs_state"], p, case["kmodes"], d)
    sc = np.abs(case["rhs_f"]).max(axis=1, keepdims=True)
    assert (np.abs(f - case["rhs_f"]) / sc).max() < 1e-14


@pytest.mark.parametrize("name", REF)
def test_oracle_step_vs_reference(name, tables):
    """Analytic Jacobian + LAPACK LU (oracle) against jacfwd + LAPACK LU (reference): cond(W) eps apart."""
    case, p, d = _setup(name, tables)
    y1, err = O.rodas5_step(case["step_t0"], case["step_t1"], case["rhs_state"], p, case["kmodes"], d)
    sc = np.abs(case["step_y1"]).max(axis=1, keepdims=True)
    assert (np.abs(y1 - case["step_y1"]) / sc).max() < 1e-8      # (scrambled states: cond(W) eps reaches 2e-9)
    sce = np.abs(case["step_err"]).max(axis=1, keepdims=True)
    assert (np.abs(err - case["step_err"]) / sce).max() < 1e-6


@pytest.mark.parametrize("name", REF)
def test_oracle_outputs_vs_reference(name, tables):
    case, p, d = _setup(name, tables)
    ks = case["kmodes"]
    y20 = O.convert_to_output(case["yfull"], p, ks[:, None], d)
    for m in range(len(ks)):
        assert helpers.field_scaled_diff(y20[m], case["y"][m]).max() < 1e-12
    pk = O.get_power(k=ks[:, None], y=case["y"], idx=4, param=p)
    np.testing.assert_allclose(pk, case["pk4"], rtol=1e-13)


@pytest.mark.parametrize("name", [n for n in REF if n[4:] in helpers.CASES])
def test_oracle_solve_vs_reference(name, tables):
    """Whole free-running solves: the oracle's committed vectors (oracle_<case>.npz) against the reference's.  Both
    step sequences start identical and stay so for every mode of at most ~150 steps; afterwards the controller's
    sensitivity to round-off lets accept/reject decisions differ, and the solutions then differ by
    the solver's own error, O(10 rtol)."""
    ref, ora = helpers.load_case(name), helpers.load_case(name[4:])
    rtol = float(ref["rtol"])
    same = ref["nsteps"] == ora["nsteps"]
    assert same.mean() >= 0.5
    for m in range(len(ref["kmodes"])):
        # (the error estimate k_8 is a small difference of large terms: at k = 10/Mpc the two LU solves already differ by
        #  3e-6 in E at the very first step, which the PID law feeds into dt)
        ns = int(min(ref["nsteps"][m], ora["nsteps"][m], 10))
        np.testing.assert_allclose(ora["rp_tnext"][m, :ns], ref["rp_tnext"][m, :ns], rtol=2e-5)
        assert np.array_equal(ora["rp_keep"][m, :ns], ref["rp_keep"][m, :ns])
        if same[m] and ref["nsteps"][m] <= 100:
            assert helpers.field_scaled_diff(ora["y"][m], ref["y"][m]).max() < 1e-6, (name, m)
        rel = np.abs(ora["y"][m][:, pc.MATTER_FIELDS] / ref["y"][m][:, pc.MATTER_FIELDS] - 1).max()
        assert rel < 50 * rtol, (name, m, rel)


@pytest.mark.parametrize("name", REF)
def test_kernel_source_step_vs_reference(emu_lib, tables, name):
    pc.check_reference_step(emu_lib, tables, name)


@pytest.mark.parametrize("name", REF)
def test_kernel_source_replays_reference_step_sequence(emu_lib, tables, name):
    """The kernel source follows the REFERENCE's own accepted/rejected step sequence and must land on the REFERENCE's
    outputs (20 fields and raw state) to 1e-6 of each field's magnitude (north_star: 1e-5)."""
    pc.check_replay(emu_lib, tables, name)


@pytest.mark.parametrize("name", REF)
def test_kernel_source_adaptive_vs_reference(emu_lib, tables, name):
    pc.check_adaptive(emu_lib, tables, name)


@pytest.mark.parametrize("name", REF)
def test_lane_kernel_source_replays_reference_step_sequence(emu_lib, tables, name, monkeypatch):
    monkeypatch.setenv("DEB_EMU_LANE", "1")
    pc.check_reference_step(emu_lib, tables, name)
    pc.check_replay(emu_lib, tables, name)


# ---- forward tangents: the complex-step oracle against the differenced reference (tools/make_reference_tangent.py) ----
@pytest.mark.parametrize("name", pc.REFERENCE_TANGENT_CASES)
def test_tangent_oracle_vs_differenced_reference(name):
    case = pc.load_tangent_case(name)
    ref = pc.load_reference_tangent(name)
    worst = pc.reference_tangent_diffs(name, case["dy"], case["dyfull"], case["dpk4"])
    # the pieces in front of the solve: initial conditions, start time, output times
    for im, m in enumerate(ref["modes"]):
        for idd, d in enumerate(ref["dir_index"]):
            s0 = np.abs(case["dy0"][d, m]).max()
            assert np.all(np.abs(ref["dy0"][im, idd] - case["dy0"][d, m]) <= 1e-7 * s0 + 10 * ref["dy0_err"][im, idd]), (name, m, d)
            if case["dtau_start"][d, m] != 0.0:
                assert abs(ref["dtau_start"][im, idd] / case["dtau_start"][d, m] - 1) < 1e-6, (name, m, d)
                np.testing.assert_allclose(ref["dtau_out"][im, idd], case["dtau_out"][d], rtol=1e-7)
    print(name, "worst scaled deviation on the well-resolved fields:", worst)


@pytest.mark.parametrize("name", pc.REFERENCE_TANGENT_CASES)
def test_kernel_source_tangent_vs_differenced_reference(emu_lib, name):
    pc.check_tangent_replay_vs_reference(emu_lib, name)
