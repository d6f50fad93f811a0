"""GPU suite (pytest -m gpu, on a B200): the CUDA library through its C-ABI against the oracle's
committed golden vectors and against the oracle itself.  Tolerances are stated in
tests/parity_checks.py.  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest

import helpers
import parity_checks as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", helpers.CASES)
def test_prologue_matches_oracle(gpu_lib, tables, name):
    pc.check_prologue(gpu_lib, tables, name)


@pytest.mark.parametrize("name", helpers.CASES)
def test_single_step_matches_dense_lu_oracle(gpu_lib, tables, name):
    pc.check_single_step(gpu_lib, tables, name)


@pytest.mark.parametrize("name", helpers.CASES)
def test_replay_of_oracle_step_sequence(gpu_lib, tables, name):
    pc.check_replay(gpu_lib, tables, name)


@pytest.mark.parametrize("name", helpers.CASES)
def test_adaptive_solve_against_oracle(gpu_lib, tables, name):
    pc.check_adaptive(gpu_lib, tables, name)


@pytest.mark.parametrize("variant", ["default", "warp", "team", "lane"])
@pytest.mark.parametrize("name", helpers.ref_cases())
def test_against_the_reference_itself(gpu_lib, tables, name, variant, monkeypatch):
    """tests/golden/ref_*.npz were produced by the REFERENCE's own sources (tools/make_reference_fixtures.py).  Every
    kernel variant takes one Rodas5 step from the reference's inputs (1e-7 of the state norm), replays the
    reference's accepted/rejected step sequence onto the reference's outputs (1e-6 of each field, raw state and 20
    fields) and solves free-running (equal step counts and 1e-5 on short modes, 50 rtol on the matter fields)."""
    if variant != "default":
        monkeypatch.setenv("DEB_VARIANT", variant)
    pc.check_reference_step(gpu_lib, tables, name)
    pc.check_replay(gpu_lib, tables, name)
    pc.check_adaptive(gpu_lib, tables, name)


@pytest.mark.parametrize("variant", ["default", "warp", "team", "lane"])
@pytest.mark.parametrize("name", helpers.BASELINE_CASES)
def test_baseline_configs_at_their_own_size(gpu_lib, tables, name, variant, monkeypatch):
    """BASELINE config 2 (every 8th of the 512 modes: 64 modes, k up to 10/Mpc, the ~600-step modes
    included), config 3 (w0wa, 32 modes of the 4096-mode grid), config 4 (three default_rng(0) cosmologies x 16 k on the
    REFERENCE's own evolve_background tables), all at n = 265: replay of the oracle's step sequence at 1e-6 on the 20
    fields and the raw state, then the free-running solve, for every kernel variant."""
    if variant != "default":
        monkeypatch.setenv("DEB_VARIANT", variant)
    pc.check_replay(gpu_lib, tables, name)
    pc.check_adaptive(gpu_lib, tables, name)


@pytest.mark.parametrize("variant", ["default", "lane"])
def test_full_grid_parity_and_convergence(gpu_lib, tables, variant, monkeypatch):
    """All 512 modes of the bench workload against the oracle's committed P(k) (what the bench line's `parity` block
    reports), and the rtol ladder against the oracle's rtol = 1e-8 truth."""
    if variant != "default":
        monkeypatch.setenv("DEB_VARIANT", variant)
    rep = pc.check_full_grid_parity(gpu_lib, tables)
    assert rep["frac_within_bar"] > 0.9, rep
    conv = pc.check_convergence(gpu_lib, tables)
    print(variant, "full grid:", rep, "convergence:", conv)


@pytest.mark.parametrize("variant", ["warp", "team", "lane"])
@pytest.mark.parametrize("name", helpers.CASES)
def test_forced_kernel_variants(gpu_lib, tables, name, variant, monkeypatch):
    """Small launches pick the CTA-per-mode kernel (deb_team.cu) on their own; DEB_VARIANT pins the choice so that the
    one-warp kernel (what large launches run) and the team kernel both face the oracle in every debug mode."""
    monkeypatch.setenv("DEB_VARIANT", variant)
    pc.check_prologue(gpu_lib, tables, name)
    pc.check_single_step(gpu_lib, tables, name)
    pc.check_replay(gpu_lib, tables, name)
    pc.check_adaptive(gpu_lib, tables, name)


def test_team_kernel_full_size(gpu_lib, emu_lib, tables, monkeypatch):
    """BASELINE config 2 at full size through the CTA-per-mode kernel: (i) deterministic -- two launches give the same
    bits, and a permuted k order the same per-mode bits (a race or a missing barrier between the team's warps would
    show here); (ii) against the main+helper kernel: same step counts on most modes, delta_m within 50 rtol on all;
    (iii) against the CPU build of the same team source: median deviation at round-off level."""
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    nk = 512
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth,
                           nnu=tab.nnu, max_steps=2048, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    run = lambda k: gpu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], k, np.array([1.0]), want_pk=True)
    monkeypatch.setenv("DEB_VARIANT", "helper")
    two = run(ks)
    monkeypatch.setenv("DEB_VARIANT", "team")
    a, b = run(ks), run(ks)
    assert np.all(a["status"] == 0)
    assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["nsteps"], b["nsteps"]) and np.array_equal(a["pk"], b["pk"])
    perm = np.random.default_rng(5).permutation(nk)
    c = run(ks[perm])
    assert np.array_equal(c["y"][0], a["y"][0][perm]) and np.array_equal(c["nsteps"][0], a["nsteps"][0][perm])
    assert (a["nsteps"] == two["nsteps"]).mean() > 0.5
    assert np.abs(a["y"][0, :, 0, 4] / two["y"][0, :, 0, 4] - 1).max() < 50 * 1e-4
    monkeypatch.setenv("DEB_EMU_TEAM", "4")
    ref = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]))
    rel = np.abs(a["y"][0, :, 0, 4] / ref["y"][0, :, 0, 4] - 1)
    assert np.median(rel) < 1e-9 and rel.max() < 50 * 1e-4
    print("kernel_ms helper", two["kernel_ms"], "team", a["kernel_ms"])


def test_context_reuse_across_shapes(gpu_lib, tables, monkeypatch):
    """One cached context, alternating call shapes: the learned work list of one shape must never be
    used -- or mis-read -- by another.  num_k 100 / 101 / 100 moved the workspace inside the old arena layout by
    exactly the size of a list entry block; every call is compared with a fresh-process-equivalent run (DEB_NO_ORDER)
    and must be bit-identical, with every mode processed (status 0, never the sentinel 3)."""
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    monkeypatch.setenv("DEB_VARIANT", "team")
    ctrl = _cabi.make_ctrl(rtol=1e-3, atol=1e-3)
    def run(nk, ncosmo=1):
        dims = _cabi.make_dims(ncosmo=ncosmo, nk=nk, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth,
                               nnu=tab.nnu, max_steps=2048, power_idx=4)
        ks = np.geomspace(1e-3, 0.3, nk)
        sc = np.repeat(tab.scalars[None], ncosmo, 0); tb = np.repeat(tab.tables[None], ncosmo, 0)
        return gpu_lib.evolve_host(dims, ctrl, sc, tb, ks, np.array([1.0]), want_pk=True)
    monkeypatch.setenv("DEB_NO_ORDER", "1")
    ref = {key: run(*key) for key in ((100, 1), (101, 1), (50, 2), (37, 3))}
    monkeypatch.delenv("DEB_NO_ORDER")
    for key in ((100, 1), (101, 1), (100, 1), (50, 2), (100, 1), (37, 3), (101, 1), (100, 1), (100, 1)):
        out = run(*key)
        assert np.all(out["status"] == 0), (key, np.unique(out["status"]))
        assert np.array_equal(out["y"], ref[key]["y"]) and np.array_equal(out["nsteps"], ref[key]["nsteps"]), key


def test_two_team_cta_is_bit_identical_to_one_team_cta(gpu_lib, tables, monkeypatch):
    """k_evolve_duo (two teams per CTA, serial warps in lock-step; the default for launches of >= 2 modes per SM slot)
    against k_evolve_team: the rendezvous moves no data, so every output bit and every step count must agree -- odd
    mode counts (one team of the last CTA idle), several cosmologies, several output times, with and without a learned
    work list, every lock-step spacing."""
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    monkeypatch.setenv("DEB_VARIANT", "team")
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    def run(nk, ncosmo, dims5, nout):
        lg, lp, lr, ln, nq = dims5
        dims = _cabi.make_dims(ncosmo=ncosmo, nk=nk, nout=nout, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth,
                               nnu=tab.nnu, max_steps=2048, power_idx=4)
        ks = np.geomspace(1e-3, 1.0, nk)
        sc = np.repeat(tab.scalars[None], ncosmo, 0); tb = np.repeat(tab.tables[None], ncosmo, 0)
        return gpu_lib.evolve_host(dims, ctrl, sc, tb, ks, np.geomspace(0.05, 1.0, nout), want_pk=True)
    for shape in ((37, 1, (31, 31, 31, 31, 5), 1), (12, 3, (11, 11, 11, 8, 3), 3), (5, 1, (5, 4, 6, 3, 4), 2), (64, 1, (16, 16, 16, 16, 3), 1)):
        monkeypatch.setenv("DEB_DUO", "0")
        ref = run(*shape)
        assert np.all(ref["status"] == 0)
        for ls in ("2", "0", "1", "4", "8"):
            monkeypatch.setenv("DEB_DUO", "2"); monkeypatch.setenv("DEB_DUO_LOCKSTEP", ls)
            for rep in range(2):          # second call: the learned work list is in use
                out = run(*shape)
                assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["nsteps"], ref["nsteps"]), (shape, ls, rep)
                assert np.array_equal(out["y"], ref["y"]) and np.array_equal(out["pk"], ref["pk"]), (shape, ls, rep)


@pytest.mark.parametrize("dims5,nk", [((11, 11, 11, 8, 3), 600), ((16, 16, 16, 16, 3), 450), ((5, 4, 6, 3, 4), 500)])
def test_deep_launch_of_small_hierarchies_takes_the_lane_kernel(gpu_lib, tables, monkeypatch, dims5, nk):
    """Beyond the team kernel's range (7 modes per SM at n <= 128) the library takes the chain-lane kernel at every n since
    round 2: three cosmologies x nk modes by the default choice against the cyclic one-warp kernel (free-running: identical
    step counts on most modes, solver-tolerance agreement on the rest), every mode processed."""
    from discoeb_b200 import _cabi
    lg, lp, lr, ln, nq = dims5
    tabs = [tables["fiducial"], tables["w0wa"], tables["fiducial"]]
    sc = np.stack([t.scalars for t in tabs]); tb = np.stack([t.tables for t in tabs])
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    dims = _cabi.make_dims(ncosmo=3, nk=nk, nout=2, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tabs[0].nth, nnu=tabs[0].nnu,
                           max_steps=4096, power_idx=4)
    ks = np.geomspace(1e-4, 5.0, nk)
    a = gpu_lib.evolve_host(dims, ctrl, sc, tb, ks, np.array([0.3, 1.0]), want_pk=True)
    monkeypatch.setenv("DEB_VARIANT", "warp")
    b = gpu_lib.evolve_host(dims, ctrl, sc, tb, ks, np.array([0.3, 1.0]), want_pk=True)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    rel = np.abs(a["pk"] / b["pk"] - 1)
    # (P(k) ~ delta^2; at lmax = 11 and rtol 1e-4 ANY two kernels differ by up to ~2e-2 on the worst long mode -- team vs
    #  cyclic 1.69e-2, lane vs cyclic 1.68e-2 on this launch, all three equally far from a tight-tolerance run:
    #  tools/compare_variants_small_n.py -- so the bar is the variants' own spread, and most modes must agree far better)
    assert np.median(rel) < 1e-8 and np.mean(rel > 1e-3) < 0.05 and rel.max() < 0.05, (np.median(rel), np.mean(rel > 1e-3), rel.max())
    assert np.mean(a["nsteps"] == b["nsteps"]) > 0.5
    assert np.array_equal(a["pk"][0], a["pk"][2]) and not np.array_equal(a["pk"][0], a["pk"][1])      # cosmologies kept apart


def test_sharded_host_entry_keeps_its_arena_across_shapes(gpu_lib, tables, monkeypatch):
    """deb_evolve_sharded_host_f64 keeps its device arena, stream and learned work list on the communicator: alternating
    call shapes (growing and shrinking) on one communicator must reproduce the plain host entry bit for bit."""
    from discoeb_b200 import _cabi
    from discoeb_b200.distributed import NativeComm
    tab = tables["fiducial"]
    monkeypatch.setenv("DEB_VARIANT", "team")
    ctrl = _cabi.make_ctrl(rtol=1e-3, atol=1e-3)
    comm = NativeComm(1, 0, lambda ident: ident, device=0, lib=gpu_lib)
    try:
        for nk, ncosmo in ((40, 1), (41, 1), (40, 1), (300, 1), (20, 2), (300, 1), (40, 1)):
            dims = _cabi.make_dims(ncosmo=ncosmo, nk=nk, nout=2, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth,
                                   nnu=tab.nnu, max_steps=2048, power_idx=4)
            ks = np.geomspace(1e-3, 0.3, nk)
            sc = np.repeat(tab.scalars[None], ncosmo, 0); tb = np.repeat(tab.tables[None], ncosmo, 0)
            a = gpu_lib.evolve_sharded_host(comm, dims, ctrl, sc, tb, ks, np.array([0.5, 1.0]), want_pk=True)
            b = gpu_lib.evolve_host(dims, ctrl, sc, tb, ks, np.array([0.5, 1.0]), want_pk=True)
            assert np.all(a["status"] == 0), (nk, ncosmo)
            for key in ("y", "pk", "tau_out", "nsteps"):
                assert np.array_equal(a[key], b[key]), (nk, ncosmo, key)
    finally:
        comm.close()


def test_gpu_matches_cpu_build_of_same_source(gpu_lib, emu_lib, tables):
    """Same source, two compilers: any difference beyond round-off is a GPU-only defect
    (missing __syncwarp, shuffle misuse, shared-memory race)."""
    case = helpers.load_case("config2_n265")
    tab = tables["fiducial"]
    ks, aout = case["kmodes"], case["aexp_out"]
    from discoeb_b200 import _cabi
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    dims = pc.dims_for(case, tab, len(ks), len(aout), return_full=True)
    a, na = gpu_lib.debug_replay(dims, ctrl, tab.scalars, tab.tables, ks, aout, case["rp_tnext"], case["rp_keep"], case["nsteps"])
    b, nb = emu_lib.debug_replay(dims, ctrl, tab.scalars, tab.tables, ks, aout, case["rp_tnext"], case["rp_keep"], case["nsteps"])
    for m in range(len(ks)):
        assert helpers.field_scaled_diff(a[0, m], b[0, m]).max() < 1e-6


def test_class_golden_curve_full_size(gpu_lib, tables):
    """Reference acceptance test at its full size (512 modes, lmax=31, nq=5, z=99):
    P_bc within 0.5 % of CLASS for k <= 10/Mpc (tests/test_perturbations.py:95-109)."""
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    g = json.load(open(os.path.join(helpers.GOLD, "CLASS_data.json")))
    kc, Pc = np.array(g["k"]), np.array(g["Pkbc"])
    nk = 512
    ks = np.geomspace(1e-5, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth,
                           nnu=tab.nnu, max_steps=2048, power_idx=6)
    out = gpu_lib.evolve_host(dims, _cabi.make_ctrl(rtol=1e-4, atol=1e-4), tab.scalars[None], tab.tables[None], ks,
                              np.array([0.01]), want_pk=True)
    assert np.all(out["status"] == 0)
    m = (kc >= 1e-5) & (kc <= 10.0)
    np.testing.assert_allclose(np.interp(kc[m], ks, out["pk"][0, :, 0]), Pc[m], rtol=0.005)


def test_full_size_properties_config2(gpu_lib, emu_lib, tables):
    """BASELINE config 2 at full size (512 modes, n=265, z=0): size-independent properties.
    (i) every mode completes; (ii) results do not depend on which warp/SM ran the mode: a
    permuted k order gives bit-identical per-mode output; (iii) linearity of the epilogue:
    pk == get_power(y); (iv) the statistical agreement with the CPU build of the same source:
    median relative deviation of delta_m at round-off level, every mode within 50 rtol."""
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    nk = 512
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth,
                           nnu=tab.nnu, max_steps=2048, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    out = gpu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]), want_pk=True)
    assert np.all(out["status"] == 0)
    perm = np.random.default_rng(3).permutation(nk)
    outp = gpu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks[perm], np.array([1.0]), want_pk=True)
    assert np.array_equal(outp["y"][0], out["y"][0][perm])
    assert np.array_equal(outp["nsteps"][0], out["nsteps"][0][perm])
    p = tab.param()
    Pk = 2 * np.pi ** 2 * p["A_s"] * (ks / p["k_p"]) ** (p["n_s"] - 1) * ks ** (-3) * out["y"][0, :, 0, 4] ** 2
    np.testing.assert_allclose(out["pk"][0, :, 0], Pk, rtol=1e-13)
    ref = emu_lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.array([1.0]))
    rel = np.abs(out["y"][0, :, 0, 4] / ref["y"][0, :, 0, 4] - 1)
    assert np.median(rel) < 1e-9
    assert rel.max() < 50 * 1e-4


def test_full_size_properties_config3(gpu_lib, tables, monkeypatch):
    """BASELINE config 3 at full size (w0wa + massive nu, n=265, 4096 modes): every mode completes, and a mode's result
    does not depend on what else is in the launch -- every 8th mode re-run as a 512-mode launch of the same (pinned)
    kernel gives the same bits, step counts included; the automatic choice (team kernel for the small launch) agrees to
    the free-running tolerance."""
    from discoeb_b200 import _cabi
    tab = tables["w0wa"]
    nk = 4096
    ks = np.geomspace(1e-4, 10.0, nk)
    mk = lambda n_: _cabi.make_dims(ncosmo=1, nk=n_, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth,
                                    nnu=tab.nnu, max_steps=2048, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    run = lambda k: gpu_lib.evolve_host(mk(len(k)), ctrl, tab.scalars[None], tab.tables[None], k, np.array([1.0]), want_pk=True)
    monkeypatch.setenv("DEB_VARIANT", "warp")
    big = run(ks)
    assert np.all(big["status"] == 0) and np.all(np.isfinite(big["pk"])) and np.all(big["pk"] > 0)
    sub = run(ks[::8])
    assert np.array_equal(sub["nsteps"][0], big["nsteps"][0][::8])
    assert np.array_equal(sub["y"][0], big["y"][0][::8]) and np.array_equal(sub["pk"][0], big["pk"][0][::8])
    monkeypatch.delenv("DEB_VARIANT")
    auto = run(ks[::8])
    assert np.all(auto["status"] == 0)
    assert np.abs(auto["y"][0, :, 0, 4] / big["y"][0, ::8, 0, 4] - 1).max() < 50 * 1e-4


def test_tight_tolerance_meets_1e5(gpu_lib, tables):
    """At rtol=atol=1e-7 the free-running solve agrees with the oracle within the 1e-5 of
    north_star on the matter transfer functions, whatever the step sequences do."""
    from discoeb_b200 import _cabi
    import oracle.discoeb_oracle as O
    tab = tables["fiducial"]
    p = tab.param()
    ks = np.array([1e-3, 0.05, 0.5])
    y, k, _ = O.evolve_perturbations(param=p, aexp_out=[1.0], kmin=0, kmax=0, num_k=3, kmodes=ks, rtol=1e-7, atol=1e-7,
                                     max_steps=20000)
    dims = _cabi.make_dims(ncosmo=1, nk=3, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth,
                           nnu=tab.nnu, max_steps=20000)
    out = gpu_lib.evolve_host(dims, _cabi.make_ctrl(rtol=1e-7, atol=1e-7), tab.scalars[None], tab.tables[None], ks, np.array([1.0]))
    assert np.all(out["status"] == 0)
    rel = np.abs(out["y"][0, :, 0, :][:, pc.MATTER_FIELDS] / y[:, 0, :][:, pc.MATTER_FIELDS] - 1)
    assert rel.max() < 1e-5


def test_host_api_drop_in(gpu_lib, tables):
    """The reference-shaped Python API: 3-tuple, param side effects, get_power, failure mode."""
    from discoeb_b200.perturbations import evolve_perturbations, evolve_perturbations_batched, get_power, MaxStepsReached
    case = helpers.load_case("default_n72")
    p = tables["fiducial"].param()
    y, k, pout = evolve_perturbations(param=p, aexp_out=case["aexp_out"], kmin=1e-4, kmax=10.0, num_k=8)
    assert y.shape == (8, 3, 20) and pout is p
    np.testing.assert_allclose(k, case["kmodes"], rtol=1e-15)
    assert (p["lmaxg"], p["lmaxgp"], p["lmaxr"], p["lmaxnu"], p["nqmax"], p["nout"]) == (11, 11, 11, 8, 3, 3)
    np.testing.assert_allclose(p["tau_out"], case["tau_out"], rtol=1e-14)
    rel = np.abs(y[..., pc.MATTER_FIELDS] / case["y"][..., pc.MATTER_FIELDS] - 1)
    assert rel.max() < 50 * 1e-4
    Pk = get_power(k=k, y=y[:, -1, :], idx=4, param=p)
    assert Pk.shape == (8,) and np.all(Pk > 0)
    yf, _, _ = evolve_perturbations(param=p, aexp_out=[1.0], kmin=1e-3, kmax=1.0, num_k=4, return_full=True)
    assert yf.shape == (4, 1, 72)
    yb, kb = evolve_perturbations_batched(param=p, aexp_out=[1.0], kmin=1e-3, kmax=1.0, num_k=4, batch_size=2)
    assert yb.shape == (4, 1, 20)
    with pytest.raises(ValueError):
        evolve_perturbations_batched(param=p, aexp_out=[1.0], kmin=1e-3, kmax=1.0, num_k=5, batch_size=2)
    with pytest.raises(MaxStepsReached):
        evolve_perturbations(param=p, aexp_out=[1.0], kmin=1e-3, kmax=1.0, num_k=4, max_steps=30)


def test_batch_of_cosmologies(gpu_lib, tables):
    from discoeb_b200.perturbations import evolve_perturbations, evolve_perturbations_multi
    ps = [tables["fiducial"].param(), tables["w0wa"].param(), tables["massless"].param()]
    y, k, info = evolve_perturbations_multi(params=ps, aexp_out=[0.5, 1.0], kmin=1e-3, kmax=1.0, num_k=16, power_idx=4)
    assert y.shape == (3, 16, 2, 20) and info["pk"].shape == (3, 16, 2)
    for i, p in enumerate(ps):
        yi, _, _ = evolve_perturbations(param=p, aexp_out=[0.5, 1.0], kmin=1e-3, kmax=1.0, num_k=16)
        assert np.array_equal(yi, y[i])


def test_edge_shapes(gpu_lib, tables):
    pc.check_edge_shapes(gpu_lib, tables)


@pytest.mark.parametrize("name", pc.BATCHED_CASES)
def test_batched_shared_step_replay(gpu_lib, tables, name):
    pc.check_batched_replay(gpu_lib, tables, name)


@pytest.mark.parametrize("name", pc.BATCHED_CASES)
def test_batched_shared_step_adaptive(gpu_lib, tables, name):
    pc.check_batched_adaptive(gpu_lib, tables, name)


def test_python_batched_api_matches_batched_oracle(gpu_lib, tables):
    """evolve_perturbations_batched = the reference's shared-step numerics (row B): the low-k batch reproduces the
    batched oracle to 1e-5 and differs from the per-mode solve at O(rtol), as the reference's two entry points do."""
    from discoeb_b200.perturbations import evolve_perturbations_batched, evolve_perturbations
    case = helpers.load_case("batched_lowk_n72")
    p = tables[str(case["cosmology"])].param()
    kw = dict(aexp_out=case["aexp_out"], kmin=1e-4, kmax=3e-3, num_k=8)
    yb, kb = evolve_perturbations_batched(param=p, batch_size=8, **kw)
    np.testing.assert_allclose(kb, case["kmodes"], rtol=1e-14)
    for m in range(8):
        assert helpers.field_scaled_diff(yb[m], case["y"][m]).max() < 1e-5
    yu, _, _ = evolve_perturbations(param=p, **kw)
    yn, _ = evolve_perturbations_batched(param=p, batch_size=8, shared_step=False, **kw)
    assert np.array_equal(yn, yu)
    d = np.abs(yb[..., 4] / yu[..., 4] - 1).max()
    assert 1e-9 < d < 50e-4


@pytest.mark.parametrize("batch_size", [4, 8, 16, 32, 64])
def test_class_golden_curve_batched(gpu_lib, tables, batch_size):
    """The reference's acceptance tests of the batched variant at full size (tests/test_perturbations.py:95-125:
    512 modes, lmax=31, nq=5, z=99, rtol=atol=1e-4, batch sizes 4...64): P_bc within 0.5 % of CLASS for k <= 10/Mpc."""
    from discoeb_b200 import _cabi
    tab = tables["fiducial"]
    g = json.load(open(os.path.join(helpers.GOLD, "CLASS_data.json")))
    kc, Pc = np.array(g["k"]), np.array(g["Pkbc"])
    nk = 512
    ks = np.geomspace(1e-5, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth,
                           nnu=tab.nnu, max_steps=2048, power_idx=6, batch_size=batch_size)
    out = gpu_lib.evolve_host(dims, _cabi.make_ctrl(rtol=1e-4, atol=1e-4), tab.scalars[None], tab.tables[None], ks,
                              np.array([0.01]), want_pk=True)
    assert np.all(out["status"] == 0)
    ns = out["nsteps"][0].reshape(-1, batch_size)
    assert np.all(ns == ns[:, :1])
    m = (kc >= 1e-5) & (kc <= 10.0)
    np.testing.assert_allclose(np.interp(kc[m], ks, out["pk"][0, :, 0]), Pc[m], rtol=0.005)
    print("batch", batch_size, "kernel_ms", out["kernel_ms"], "steps per batch", ns[:, 0].min(), "...", ns[:, 0].max())


def test_spectra_epilogues_on_gpu(gpu_lib):
    """power_multipoles, power_Kaiser, get_power_smoothed, get_xi_from_P computed by the CUDA library
    (csrc/deb_spectra.cu) against the NumPy oracle (oracle/spectra.py, itself pinned against SciPy)."""
    from test_spectra_cpu import check_product_spectra
    check_product_spectra(gpu_lib)
