"""Table production: the producer source (csrc/deb_background.cuh) against the REFERENCE's own
``evolve_background`` -- tests/golden/reference_background.npz holds what the unmodified reference returned when run under
tools/refshim (tools/make_reference_background.py) for 7 cosmologies (test fiducial, w0wa, massless, Fisher fiducial,
three BASELINE config-4 draws).  CPU: the C++ build of the source (tests/emu); GPU (-m gpu): the CUDA kernel through the
C-ABI, plus the whole chain parameters -> tables -> P(k) on the device against the reference's CLASS acceptance curve.

Bars: knots 1e-13; table values 1e-6 of the table's magnitude (measured 1e-12 .. 1e-15 on the CPU build, up to 2e-7 on
the GPU for a c_s^2, which holds the difference 4 - d ln(a T_m)/d ln a and whose step sizes follow CUDA's pow/exp;
the HeII Saha branch switches formula at R = 1e5, where the two closed forms differ by ~1e-8); second derivatives 1e-4 (they amplify value noise by
1/h^2).  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest

import helpers
from discoeb_b200._pack import SPLINE_KEYS

Z = np.load(os.path.join(helpers.GOLD, "reference_background.npz"))
NAMES = [str(n) for n in Z["names"]]
NTH, NNU = 256, 512


def _check_against_reference(lib, ybar=1e-6, sbar=1e-4):
    bg_in = np.stack([Z[f"{n}_in"] for n in NAMES])
    scal, tab, _ = lib.background_host(bg_in, NTH)
    for c, n in enumerate(NAMES):
        rs = Z[f"{n}_scalars"]
        np.testing.assert_allclose(scal[c, :16], rs, rtol=1e-13, atol=1e-300)
        off = 0
        for j, key in enumerate(SPLINE_KEYS):
            m = NNU if j in (2, 3) else NTH
            x, y, S = tab[c, off:off + m], tab[c, off + m:off + 2 * m], tab[c, off + 2 * m:off + 3 * m]
            off += 3 * m
            rx, ry, rS = Z[f"{n}_{key}_x"], Z[f"{n}_{key}_y"], Z[f"{n}_{key}_S"]
            assert np.abs(x - rx).max() <= 1e-13 * np.abs(rx).max(), (n, key)
            assert np.abs(y - ry).max() <= ybar * np.abs(ry).max(), (n, key)
            assert np.abs(S - rS).max() <= sbar * np.abs(rS).max(), (n, key)
    return scal, tab


def test_producer_source_matches_reference_evolve_background(emu_lib):
    _check_against_reference(emu_lib)


ZX = np.load(os.path.join(helpers.GOLD, "reference_background_extras.npz"))


def _check_extras_against_reference(lib, bar=1e-6, sbar=1e-4, vbar=1e-6):
    """The rest of evolve_background's outputs (species fractions, c_s^2, T_m, x_e', opacity, optical depth, visibility and
    its derivatives, the two extra tau-splines, the pseudo-pressure table) against what the reference left in ``param``
    (tools/make_reference_background.py -> reference_background_extras.npz).  Every array relative to its own largest
    magnitude; the visibility derivatives and x_e' come from spline second derivatives and inherit their bar."""
    from discoeb_b200.background import EXTRA_ROWS, unpack_extras, unpack_param
    bg_in = np.stack([Z[f"{n}_in"] for n in NAMES])
    scal, tab, _, ext = lib.background_host(bg_in, NTH, extras=True)
    worst = {}
    for c, n in enumerate(NAMES):
        p = unpack_extras(unpack_param({}, scal[c], tab[c], NTH), ext[c], NTH)
        for key in EXTRA_ROWS:
            ref = ZX[f"{n}_{key}"]
            tol = {"xeprime": sbar, "xeprime_recf": vbar, "gvisprime": sbar, "gvispprime": 10 * sbar}.get(key, bar)
            got = p[key]
            if key == "xeprime_recf":
                # In the Saha-H intervals (thermodynamics_recfast.py:418-432) the reference's closed-form d x_H / dz is the
                # difference of two terms ~1e12 apart from their sum: its value there is cancellation noise (1e3 x the true
                # derivative, sign changes from knot to knot, also between the reference's own runs on different libm's).
                # Compared outside those intervals (x_H > 0.99 with helium already recombining), one knot of margin.
                sh = (ZX[f"{n}_xeHI"] > 0.99) & (ZX[f"{n}_xeHeI"] < 0.995)
                sh = sh | np.roll(sh, 1) | np.roll(sh, -1)
                assert 5 <= sh.sum() <= 40, (n, int(sh.sum()))
                got, ref = got[~sh], ref[~sh]
            # (the natural spline of the steeply falling opacity overshoots between the first knots, a << 1e-6: the reference's
            #  optical depth dips below -700 there and its visibility overflows to inf on a handful of knots; so does ours)
            fin = np.isfinite(ref)
            assert fin.sum() >= len(ref) - 8, (n, key)
            e = np.abs(got[fin] - ref[fin]).max() / np.abs(ref[fin]).max()
            worst[key] = max(worst.get(key, 0.0), e)
            assert e <= tol, (n, key, e)
        for key in ("cs2a_of_tau_spline", "tempba_of_tau_spline", "logppseudonu_of_loga_spline"):
            for part, tol in (("y", bar), ("S", sbar)):
                ref = ZX[f"{n}_{key}_{part}"]
                e = np.abs(getattr(p[key], part) - ref).max() / np.abs(ref).max()
                worst[key + "." + part] = max(worst.get(key + "." + part, 0.0), e)
                assert e <= tol, (n, key, part, e)
        for key in ("taumax", "adotrad", "Omegamnu"):
            assert abs(p[key] / float(ZX[f"{n}_{key}"]) - 1) < 1e-12 or float(ZX[f"{n}_{key}"]) == 0.0, (n, key)
        # the visibility function integrates to one over the history (a property, whatever the reference)
        late = p["aexp"] > 5e-4          # (z < 2000: the early knots ring, see above; the visibility peak sits at z ~ 1100)
        assert abs(np.trapezoid(p["gvis"][late], p["tau"][late]) - 1.0) < 0.01, n
    return worst


def test_producer_extras_match_reference_evolve_background(emu_lib):
    worst = _check_extras_against_reference(emu_lib)
    print({k: float(f"{v:.1e}") for k, v in worst.items()})


def test_table_spline_members_match_reference_spline():
    """derivative / derivative2 / integral of the host-side spline object against the relations the reference's arrays obey:
    optical_depth = integral of the opacity spline from tau to the end, minus today's value; gvisprime from derivative12."""
    from discoeb_b200.background import TableSpline
    n = NAMES[0]
    tau, opac = Z[f"{n}_xe_of_tau_spline_x"], ZX[f"{n}_opac"]
    import oracle.background as OB
    sp = OB.Spline(tau, opac)
    ts = TableSpline(tau, opac, sp.S)
    ts.integrate_from_start = False
    tau0 = TableSpline(Z[f"{n}_tau_of_a_spline_x"], Z[f"{n}_tau_of_a_spline_y"], Z[f"{n}_tau_of_a_spline_S"]).evaluate(1.0)
    od = ts.integral(tau) - ts.integral(tau0)
    ref = ZX[f"{n}_optical_depth"]
    assert np.abs(od - ref).max() <= 1e-10 * np.abs(ref).max()
    d1, d2 = ts.derivative12(tau)
    late = Z[f"{n}_tau_of_a_spline_x"] > 5e-4          # (the early knots overflow in the reference itself, see above)
    em = np.exp(-ref[late])
    g1, g2 = ZX[f"{n}_gvisprime"][late], ZX[f"{n}_gvispprime"][late]
    assert np.abs((d1[late] + opac[late] ** 2) * em - g1).max() <= 1e-9 * np.abs(g1).max()
    assert np.abs((d2[late] + 3 * opac[late] * d1[late] + opac[late] ** 3) * em - g2).max() <= 1e-9 * np.abs(g2).max()


def test_reference_recfast_curve_status(emu_lib):
    """The reference's own regression curve (tests/resources/RECFAST_DISCO_EB_data.json, "v0.1.0 baseline", 0.5 % in
    tests/test_background.py:60-75) is NOT met by the reference's current algorithm as executed here (GRKT4 at rtol
    1e-3 on 256 adaptive knots: up to 7 % low around z ~ 1000, the same with the reference's sources under refshim), nor by
    a converged LSODA solution of the same equations (5 %).  The producer is held to the reference's algorithm (test
    above); this test records the distance to the old curve so that a change shows up."""
    g = json.load(open(os.path.join(helpers.GOLD, "RECFAST_DISCO_EB_data.json")))
    a_g, xe_g = np.array(g["a"]), np.array(g["xe"])
    from discoeb_b200.background import unpack_param
    scal, tab, _ = emu_lib.background_host(Z["fiducial_in"][None], NTH)
    p = unpack_param({}, scal[0], tab[0], NTH)
    xe = p["xe_of_tau_spline"].evaluate(p["tau_of_a_spline"].evaluate(a_g))
    rel = np.abs(xe / xe_g - 1)
    assert np.median(rel) < 1e-5 and rel.max() < 0.08


def test_python_api_mirrors_reference_keys(emu_lib):
    from discoeb_b200.background import evolve_background
    p = evolve_background(param=dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09,
                                     n_s=0.96822, H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06), lib=emu_lib)
    for key in SPLINE_KEYS + ("grhom", "grhog", "grhor", "amnu", "OmegaDE", "taumin", "taumax", "Omegamnu", "aexp", "tau", "xe",
                              # everything else background.py:144-342 leaves in param
                              "amin", "amax", "adotrad", "a", "logppseudonu_of_loga_spline", "xeHI", "xeHeI", "xeHeII", "cs2", "Tm",
                              "cs2a_of_tau_spline", "tempba_of_tau_spline", "optical_depth", "opac", "gvis", "gvisprime", "gvispprime",
                              "xeprime_recf", "xeprime", "fHe"):
        assert key in p, key
    assert abs(p["tau_of_a_spline"].evaluate(1.0) / 14165.0 - 1) < 0.01          # conformal age ~ 14.2 Gpc


@pytest.mark.gpu
def test_gpu_producer_matches_reference_evolve_background(gpu_lib):
    """CUDA's exp/pow differ from glibc's in the last bit, and RECFAST's right-hand side is DISCONTINUOUS (hard switches
    at x_H = 0.99 / 0.985, x_He = 0.98 ...: thermodynamics_recfast.py:178-262): a stage evaluated within an ulp of a
    switch takes the other branch.  Measured GPU vs reference / CPU build: x_e to 1.5e-6, second derivatives to 2e-4
    (64 config-4 draws, profiles/r2_background_*); the reference on JAX-GPU vs JAX-CPU is exposed to the same."""
    _check_against_reference(gpu_lib, ybar=1e-5, sbar=1e-3)


@pytest.mark.gpu
def test_gpu_producer_extras_match_reference_evolve_background(gpu_lib):
    """Same looser bars as the tables on the GPU (discontinuous RECFAST right-hand side + CUDA libm, see above)."""
    worst = _check_extras_against_reference(gpu_lib, bar=1e-5, sbar=1e-3, vbar=1e-4)
    print({k: float(f"{v:.1e}") for k, v in worst.items()})


@pytest.mark.gpu
def test_gpu_batch_matches_cpu_build(gpu_lib, emu_lib):
    """64 BASELINE config-4 draws in one launch: same tables as the CPU build of the source, cosmology by cosmology."""
    from discoeb_b200.background import config4_draws, pack_background_input
    base = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09, n_s=0.96822,
                H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06)
    bg_in = np.stack([pack_background_input({**base, **d}) for d in config4_draws(64)])
    sg, tg, ms = gpu_lib.background_host(bg_in, NTH)
    se, te, _ = emu_lib.background_host(bg_in, NTH)
    np.testing.assert_allclose(sg, se, rtol=1e-12, atol=1e-300)
    # table by table, relative to each table's magnitude
    off = 0
    for j in range(7):
        m = NNU if j in (2, 3) else NTH
        for part in range(3):
            a, b = tg[:, off:off + m], te[:, off:off + m]
            bar = 1e-5 if part < 2 else 1e-3
            assert (np.abs(a - b).max(axis=1) <= bar * np.abs(b).max(axis=1)).all(), (j, part)
            off += m
    print(f"background: 64 cosmologies in {ms:.2f} ms on the GPU")


@pytest.mark.gpu
def test_gpu_parameters_to_pk_against_class(gpu_lib):
    """Parameters -> tables (GPU) -> P_bc(k) at z = 99 (GPU) against the reference's CLASS acceptance curve
    (tests/test_perturbations.py:95-109, 0.5 %, k <= 10/Mpc)."""
    from discoeb_b200.background import evolve_background
    from discoeb_b200.perturbations import evolve_perturbations
    p = evolve_background(param=dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09,
                                     n_s=0.96822, H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06, k_p=0.05))
    y, k, _ = evolve_perturbations(param=p, aexp_out=np.array([0.01]), kmin=1e-5, kmax=10.0, num_k=512, lmaxg=31, lmaxgp=31, lmaxr=31,
                                   lmaxnu=31, nqmax=5, rtol=1e-4, atol=1e-4)
    Pk = 2 * np.pi ** 2 * p["A_s"] * (k / p["k_p"]) ** (p["n_s"] - 1) * k ** (-3) * y[:, 0, 6] ** 2
    g = json.load(open(os.path.join(helpers.GOLD, "CLASS_data.json")))
    kc, Pc = np.array(g["k"]), np.array(g["Pkbc"])
    m = (kc >= 1e-5) & (kc <= 10.0)
    np.testing.assert_allclose(np.interp(kc[m], k, Pk), Pc[m], rtol=0.005)
