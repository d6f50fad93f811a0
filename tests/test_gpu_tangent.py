"""GPU parity tests of the forward-tangent row: the CUDA tangent kernel through the
C-ABI against the complex-step oracle's committed vectors (tests/golden/tangent_*.npz)."""
import numpy as np
import pytest

import helpers
import parity_checks as pc
from discoeb_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", pc.TANGENT_CASES)
def test_tangent_replay_of_oracle_step_sequence(gpu_lib, name):
    worst = pc.check_tangent_replay(gpu_lib, name)
    print(name, "worst scaled tangent deviation", worst)


@pytest.mark.parametrize("name", pc.REFERENCE_TANGENT_CASES)
def test_tangent_replay_against_the_differenced_reference(gpu_lib, name):
    """The CUDA tangent kernel against central differences (+ Richardson) of the REFERENCE's own step function, run from
    its sources by tools/make_reference_tangent.py (tests/golden/reference_tangent_*.npz)."""
    pc.check_tangent_replay_vs_reference(gpu_lib, name)


@pytest.mark.parametrize("name", ("default_n72", "w0wa_n43", "kscaled_n72"))
def test_tangent_adaptive_against_oracle(gpu_lib, name):
    pc.check_tangent_adaptive(gpu_lib, name)


def test_tangent_properties(gpu_lib):
    pc.check_tangent_properties(gpu_lib)


def test_tangent_properties_full_size_n265(gpu_lib):
    """Config-5 shape (n=265): linearity of the tangent map, d P/d A_s, d P/d n_s over 64 modes."""
    pc.check_tangent_properties(gpu_lib, name="fisher_n265x2", nk=64)


def test_tangent_batch_of_cosmologies(gpu_lib):
    pc.check_tangent_batch_of_cosmologies(gpu_lib)


def test_tangent_primal_half_equals_plain_solve(gpu_lib, monkeypatch):
    """The primal outputs of the tangent launch are those of the one-warp primal kernel (same source, same step
    sequence; a small primal launch would otherwise pick the CTA-per-mode kernel, whose sweeps round differently)."""
    monkeypatch.setenv("DEB_VARIANT", "warp")
    case = pc.load_tangent_case("default_n72")
    ks = np.geomspace(1e-3, 1.0, 24)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    dims_t = pc.tangent_dims(case, nk=len(ks), ntan=2, power_idx=4)
    out_t = gpu_lib.evolve_tangent_host(dims_t, ctrl, case["scalars"][None], case["tables"][None], ks, case["aexp_out"],
                                        case["d_scalars"][:2, None], case["d_tables"][:2, None], want_pk=True)
    dims_p = pc.tangent_dims(case, nk=len(ks), ntan=0, power_idx=4)
    out_p = gpu_lib.evolve_host(dims_p, ctrl, case["scalars"][None], case["tables"][None], ks, case["aexp_out"], want_pk=True)
    assert np.array_equal(out_t["nsteps"], out_p["nsteps"]) or np.mean(out_t["nsteps"] == out_p["nsteps"]) > 0.5
    same = (out_t["nsteps"] == out_p["nsteps"])[0]
    for m in np.nonzero(same)[0]:
        assert helpers.field_scaled_diff(out_t["y"][0, m], out_p["y"][0, m]).max() < 1e-6


def test_tangent_matches_cpu_build_of_same_source(gpu_lib, emu_lib):
    case = pc.load_tangent_case("w0wa_n43")
    ctrl = _cabi.make_ctrl(rtol=float(case["rtol"]), atol=float(case["rtol"]))
    dims = pc.tangent_dims(case)
    args = (dims, ctrl, case["scalars"][None], case["tables"][None], case["kmodes"], case["aexp_out"], case["d_scalars"][:, None],
            case["d_tables"][:, None], case["rp_tnext"], case["rp_dtnext"], case["rp_keep"], case["nsteps"])
    yg, dyg, _, _ = gpu_lib.debug_replay_tangent(*args)
    ye, dye, _, _ = emu_lib.debug_replay_tangent(*args)
    for m in range(len(case["kmodes"])):
        for d in range(dyg.shape[0]):
            assert pc.tangent_scaled_diff(dyg[d, 0, m], dye[d, 0, m], ye[0, m]).max() < 1e-7


def test_python_jvp_api(gpu_lib):
    from discoeb_b200.perturbations import evolve_perturbations_jvp, evolve_perturbations
    case = pc.load_tangent_case("default_n72")
    p = helpers.Tables(case["scalars"], case["tables"], case["nth"], case["nnu"]).param()
    dps = [helpers.Tables(case["d_scalars"][d], case["d_tables"][d], case["nth"], case["nnu"]).param() for d in range(2)]
    y, dy, k, info = evolve_perturbations_jvp(param=p, dparam=dps, aexp_out=[0.1, 1.0], kmin=1e-3, kmax=0.5, num_k=8, power_idx=4)
    assert y.shape == (8, 2, 20) and dy.shape == (2, 8, 2, 20) and info["dpk"].shape == (2, 8, 2)
    y1, dy1, _, info1 = evolve_perturbations_jvp(param=p, dparam=dps[1], aexp_out=[0.1, 1.0], kmin=1e-3, kmax=0.5, num_k=8, power_idx=4)
    assert dy1.shape == (8, 2, 20)
    np.testing.assert_array_equal(dy1, dy[1])
    # d P = 2 P d delta / delta for a direction without A_s / n_s / k_p seeds
    np.testing.assert_allclose(info["dpk"][0], 2 * info["pk"] * dy[0][..., 4] / y[..., 4], rtol=1e-10)
    assert "tau_out" in p and p["nout"] == 2
