"""CPU suite, part 2: the KERNEL SOURCE (disco-eb_b200/csrc/deb_core.cuh) compiled as plain C++
with the warp's lanes run by loops (tests/emu) against the oracle.  This checks the algorithm the
GPU executes -- structured Jacobian, tail sweeps, pivoted head, stage logic, controller, output
sampling -- without a GPU.  The CUDA build of the same source is checked by tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

import helpers
import parity_checks as pc


@pytest.mark.parametrize("name", helpers.CASES)
def test_prologue_matches_oracle(emu_lib, tables, name):
    pc.check_prologue(emu_lib, tables, name)


@pytest.mark.parametrize("name", helpers.CASES)
def test_single_step_matches_dense_lu_oracle(emu_lib, tables, name):
    pc.check_single_step(emu_lib, tables, name)


@pytest.mark.parametrize("name", helpers.CASES)
def test_replay_of_oracle_step_sequence(emu_lib, tables, name):
    pc.check_replay(emu_lib, tables, name)


@pytest.mark.parametrize("name", helpers.CASES)
def test_adaptive_solve_against_oracle(emu_lib, tables, name):
    pc.check_adaptive(emu_lib, tables, name)


@pytest.mark.parametrize("variant", ["cyclic", "lane"])
@pytest.mark.parametrize("name", helpers.BASELINE_CASES)
def test_baseline_configs_replay_and_adaptive(emu_lib, tables, name, variant, monkeypatch):
    """BASELINE configs 2 (64 modes, k up to 10/Mpc, the ~600-step modes included), 3 (w0wa, 32 modes) and 4 (three
    default_rng(0) cosmologies x 16 k on the reference's own evolve_background tables) at n = 265: replay of the
    oracle's step sequence at 1e-6 on all 20 fields and the raw state, and the free-running solve."""
    if variant == "lane":
        monkeypatch.setenv("DEB_EMU_LANE", "1")
    pc.check_replay(emu_lib, tables, name)
    pc.check_adaptive(emu_lib, tables, name)


def test_full_grid_parity_and_convergence_through_kernel_source(emu_lib, tables):
    rep = pc.check_full_grid_parity(emu_lib, tables)
    assert rep["frac_within_bar"] > 0.5
    conv = pc.check_convergence(emu_lib, tables)
    print("full grid", rep, "convergence (kernel-truth, oracle-truth, kernel-oracle):", conv)


def test_class_golden_curve_through_kernel_source(emu_lib, tables):
    """The reference's own acceptance test (tests/test_perturbations.py:95-109): P_bc(k) at z=99,
    lmax=31, nq=5, 512 modes, rtol=atol=1e-4, linearly interpolated onto the CLASS k grid, within
    0.5 % of tests/resources/CLASS_data.json for k <= 10/Mpc."""
    import json
    import os
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    g = json.load(open(os.path.join(helpers.GOLD, "CLASS_data.json")))
    kc, Pc = np.array(g["k"]), np.array(g["Pkbc"])
    nk = 512
    ks = np.geomspace(1e-5, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth,
                           nnu=tab.nnu, max_steps=2048, power_idx=6)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    out = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([0.01]), want_pk=True)
    assert np.all(out["status"] == 0)
    p = tab.param()
    Pk = 2 * np.pi ** 2 * p["A_s"] * (ks / p["k_p"]) ** (p["n_s"] - 1) * ks ** (-3) * out["y"][0, :, 0, 6] ** 2
    np.testing.assert_allclose(out["pk"][0, :, 0], Pk, rtol=1e-13)      # fused get_power epilogue
    m = (kc >= 1e-5) & (kc <= 10.0)
    np.testing.assert_allclose(np.interp(kc[m], ks, Pk), Pc[m], rtol=0.005)


def test_max_steps_is_reported(emu_lib, tables):
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    ks = np.array([1e-3, 1.0])
    dims = _cabi.make_dims(ncosmo=1, nk=2, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth,
                           nnu=tab.nnu, max_steps=40)
    out = emu_lib.evolve_host(dims, _cabi.make_ctrl(rtol=1e-4, atol=1e-4), tab.scalars[None], tab.tables[None], ks,
                              np.array([1.0]))
    assert list(out["status"][0]) == [0, 1]
    assert out["nsteps"][0, 1] == 40


def test_two_cosmologies_in_one_call(emu_lib, tables):
    from discoeb_b200 import _cabi
    a, b = tables["fiducial"], tables["w0wa"]
    ks = np.geomspace(1e-3, 1.0, 5)
    mk = lambda nc: _cabi.make_dims(ncosmo=nc, nk=5, nout=2, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3,
                                    nth=a.nth, nnu=a.nnu, max_steps=4096)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    aout = np.array([0.5, 1.0])
    both = emu_lib.evolve_host(mk(2), ctrl, np.stack([a.scalars, b.scalars]), np.stack([a.tables, b.tables]), ks, aout)
    one_a = emu_lib.evolve_host(mk(1), ctrl, a.scalars[None], a.tables[None], ks, aout)
    one_b = emu_lib.evolve_host(mk(1), ctrl, b.scalars[None], b.tables[None], ks, aout)
    assert np.array_equal(both["y"][0], one_a["y"][0])
    assert np.array_equal(both["y"][1], one_b["y"][0])
    assert not np.allclose(both["y"][0], both["y"][1], rtol=1e-6)


def test_per_cosmology_k_grids(emu_lib, tables):
    """k_per_cosmo=1: every cosmology of a batch brings its own k grid (what a caller that scales
    kmin/kmax by h needs)."""
    from discoeb_b200 import _cabi
    a, b = tables["fiducial"], tables["w0wa"]
    k2 = np.stack([np.geomspace(1e-3, 0.5, 4), np.geomspace(2e-3, 1.0, 4)])
    mk = lambda nc, kp: _cabi.make_dims(ncosmo=nc, nk=4, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3,
                                        nth=a.nth, nnu=a.nnu, max_steps=4096, k_per_cosmo=kp)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    both = emu_lib.evolve_host(mk(2, True), ctrl, np.stack([a.scalars, b.scalars]), np.stack([a.tables, b.tables]), k2, np.array([1.0]))
    one_b = emu_lib.evolve_host(mk(1, False), ctrl, b.scalars[None], b.tables[None], k2[1], np.array([1.0]))
    assert np.array_equal(both["y"][1], one_b["y"][0])


@pytest.mark.parametrize("name", ["default_n72", "config2_n265", "many_out_n72"])
def test_two_warp_variant_source(emu_lib, tables, name, monkeypatch):
    """The main+helper code path (the helper's evaluations run inline in the CPU build): same checks, and the
    results agree with the one-warp path to round-off on modes whose step sequences coincide."""
    from discoeb_b200 import _cabi
    case = helpers.load_case(name)
    tab = tables[str(case["cosmology"])]
    ks, aout, rtol = case["kmodes"], case["aexp_out"], float(case["rtol"])
    dims = pc.dims_for(case, tab, len(ks), len(aout))
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    one = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout)
    monkeypatch.setenv("DEB_EMU_HELPER", "1")
    pc.check_adaptive(emu_lib, tables, name)
    two = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout)
    same = (one["nsteps"] == two["nsteps"]) & (one["naccept"] == two["naccept"])
    assert same.mean() >= 0.5
    for m in np.nonzero(same[0])[0]:
        if one["nsteps"][0, m] <= 100:
            assert helpers.field_scaled_diff(two["y"][0, m], one["y"][0, m]).max() < 1e-6


@pytest.mark.parametrize("team", ["2", "4", "8"])
@pytest.mark.parametrize("name", ["default_n72", "config2_n265", "odd_dims_n43", "min_dims_n33", "many_out_n72"])
def test_team_variant_source(emu_lib, tables, name, team, monkeypatch):
    """The CTA-per-mode code path (deb_team.cuh; the team's threads run by loops in the CPU build): prologue,
    single step, replay and the free-running solve against the oracle; against the main+helper path the step counts
    agree on most modes and short modes agree to round-off (the segmented tail sweeps round differently, and a
    free-running stiff solve amplifies that into different accept/reject decisions on long modes)."""
    from discoeb_b200 import _cabi
    case = helpers.load_case(name)
    tab = tables[str(case["cosmology"])]
    ks, aout, rtol = case["kmodes"], case["aexp_out"], float(case["rtol"])
    dims = pc.dims_for(case, tab, len(ks), len(aout))
    ctrl = _cabi.make_ctrl(rtol=rtol, atol=rtol)
    monkeypatch.setenv("DEB_EMU_HELPER", "1")
    two = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout, want_pk=False)
    monkeypatch.setenv("DEB_EMU_TEAM", team)
    pc.check_prologue(emu_lib, tables, name)
    pc.check_single_step(emu_lib, tables, name)
    pc.check_replay(emu_lib, tables, name)
    pc.check_adaptive(emu_lib, tables, name)
    tm = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, aout, want_pk=False)
    same = (tm["nsteps"] == two["nsteps"]) & (tm["naccept"] == two["naccept"])
    assert same.mean() >= 0.5
    for m in np.nonzero(same[0])[0]:
        if two["nsteps"][0, m] <= 100:
            assert helpers.field_scaled_diff(tm["y"][0, m], two["y"][0, m]).max() < 1e-6


def test_team_variant_edge_shapes(emu_lib, tables, monkeypatch):
    monkeypatch.setenv("DEB_EMU_TEAM", "4")
    pc.check_edge_shapes(emu_lib, tables)


def test_edge_shapes(emu_lib, tables):
    pc.check_edge_shapes(emu_lib, tables)


@pytest.mark.parametrize("name", pc.BATCHED_CASES)
def test_batched_shared_step_replay(emu_lib, tables, name):
    pc.check_batched_replay(emu_lib, tables, name)


@pytest.mark.parametrize("name", pc.BATCHED_CASES)
def test_batched_shared_step_adaptive(emu_lib, tables, name):
    pc.check_batched_adaptive(emu_lib, tables, name)
