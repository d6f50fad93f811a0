"""Spectra post-processing entry points (perturbations.py:1063-1224).  The NumPy oracle (oracle/spectra.py) is checked
against independent SciPy implementations and closed forms; the product's kernels (csrc/deb_spectra.cuh behind
discoeb_b200/spectra.py) are checked against the oracle -- here through the CPU build of the source (tests/emu), in
tests/test_gpu_parity.py through the CUDA library."""
import numpy as np
import scipy.signal
import scipy.special

import helpers
import oracle.spectra as S
from discoeb_b200 import spectra as G
from discoeb_b200 import perturbations as P


def _case():
    case = helpers.load_case("default_n72")
    p = helpers.load_tables("fiducial").param()
    return case, p


def test_reexported_from_perturbations():
    for name in ("get_power", "get_power_smoothed", "power_Kaiser", "power_multipoles", "get_xi_from_P"):
        assert getattr(P, name) is getattr(G, name)


def test_lngamma_matches_scipy():
    z = np.concatenate([0.75 + 1j * np.linspace(0, 60, 200), -2.3 + 1j * np.linspace(0.1, 5, 20), np.array([3.0 + 0j, 0.2 + 0.1j])])
    ref = scipy.special.loggamma(z)
    got = S.lngamma_complex_e(z)
    np.testing.assert_allclose(got.real, ref.real, rtol=1e-12, atol=1e-12)
    # arg Gamma is defined modulo 2 pi
    d = (got.imag - ref.imag) / (2 * np.pi)
    np.testing.assert_allclose(d, np.round(d), atol=1e-10)


def test_savgol_matches_scipy_in_the_interior():
    rng = np.random.default_rng(0)
    y = np.cumsum(rng.normal(size=400))
    for w in (5, 11, 31):
        got = S.savgol_filter(y=y, window_length=w, polyorder=3)
        ref = scipy.signal.savgol_filter(y, w, 3, mode="constant", cval=0.0)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_power_smoothed_kaiser_multipoles():
    rng = np.random.default_rng(1)
    k = np.geomspace(1e-3, 1.0, 128)
    p = dict(A_s=2.1e-9, n_s=0.96, k_p=0.05)
    y = np.zeros((128, 20))
    y[:, 4] = 1e3 * k ** 1.5 * (1 + 0.05 * np.sin(40 * np.log(k)))
    y[:, 5] = -0.5 * y[:, 4]
    Pm = S.get_power(k=k, y=y, idx=4, param=p)
    Ps = S.get_power_smoothed(k=k, y=y, dlogk=0.5, idx=4, param=p)
    w = round(0.5 / np.log(k[1] / k[0])); w += (w + 1) % 2
    assert np.array_equal(Ps[:w // 2], Pm[:w // 2]) and np.array_equal(Ps[-(w // 2) - 1:], Pm[-(w // 2) - 1:])
    ref = np.exp(scipy.signal.savgol_filter(np.log(Pm), w, 3, mode="constant", cval=0.0))
    np.testing.assert_allclose(Ps[w // 2:-(w // 2) - 1], ref[w // 2:-(w // 2) - 1], rtol=1e-9)
    assert np.std(np.log(Ps / (k ** 0.96))[w:-w]) < np.std(np.log(Pm / (k ** 0.96))[w:-w])      # it does smooth the wiggles
    # Kaiser: multipoles are the Legendre moments of P(k, mu)
    Pkmu, mu = S.power_Kaiser(y=y, kmodes=k, bias=1.7, nmu=2001, param=p)
    P0, P2, P4 = S.power_multipoles(y=y, kmodes=k, b=1.7, param=p)
    L2 = 0.5 * (3 * mu ** 2 - 1)
    L4 = (35 * mu ** 4 - 30 * mu ** 2 + 3) / 8
    trap = lambda f: np.sum(0.5 * (f[:, 1:] + f[:, :-1]) * np.diff(mu), axis=1)
    np.testing.assert_allclose(0.5 * trap(Pkmu), P0, rtol=1e-5)
    np.testing.assert_allclose(2.5 * trap(Pkmu * L2), P2, rtol=1e-4)
    np.testing.assert_allclose(4.5 * trap(Pkmu * L4), P4, rtol=1e-3)
    Pth, mu2 = S.power_Kaiser(y=y, kmodes=k, bias=1.7, mu_sampling=False, nmu=5, param=p)
    np.testing.assert_allclose(mu2, np.cos(np.linspace(0, np.pi, 5)))
    Pks, _ = S.power_Kaiser(y=y, kmodes=k, bias=1.7, smooth_dlogk=0.5, nmu=5, param=p)
    assert Pks.shape == (128, 5) and np.all(np.isfinite(Pks))


def test_fftlog_gaussian_pair():
    """P(k) = exp(-k^2 s^2)  <->  xi(r) = exp(-r^2 / 4 s^2) / (8 pi^1.5 s^3).  The reference's discretisation
    (period taken as log(kmax/kmin) for N samples, r = 2 pi / k) is accurate to a few 1e-3; it is mirrored as is."""
    N = 2048
    k = np.geomspace(1e-4, 1e2, N)
    Pk = np.exp(-k ** 2)
    xi, r = S.get_xi_from_P(k=k, Pk=Pk, ell=0)
    exact = np.exp(-r ** 2 / 4) / (8 * np.pi ** 1.5)
    m = (r > 0.2) & (r < 4.0)
    np.testing.assert_allclose(xi[m], exact[m], rtol=1e-2)
    assert np.all(np.diff(r) > 0) and xi.shape == (N,)
    # quadrupole of a pure monopole-shaped input is a different Hankel transform: finite and of the right size
    xi2, _ = S.get_xi_from_P(k=k, Pk=Pk, ell=2)
    assert np.all(np.isfinite(xi2)) and np.abs(xi2[m]).max() < np.abs(xi[m]).max()


def check_product_spectra(lib):
    """discoeb_b200.spectra (library kernels) against the oracle on a wiggly synthetic spectrum and on a solver output."""
    k = np.geomspace(1e-3, 1.0, 128)
    p = dict(A_s=2.1e-9, n_s=0.96, k_p=0.05)
    y = np.zeros((128, 20))
    y[:, 4] = 1e3 * k ** 1.5 * (1 + 0.05 * np.sin(40 * np.log(k)))
    y[:, 5] = -0.5 * y[:, 4]
    for a, b in zip(G.power_multipoles(y=y, kmodes=k, b=1.7, param=p, lib=lib), S.power_multipoles(y=y, kmodes=k, b=1.7, param=p)):
        np.testing.assert_allclose(a, b, rtol=1e-13)
    for idx in (4, 5):
        np.testing.assert_allclose(G.get_power_smoothed(k=k, y=y, dlogk=0.5, idx=idx, param=p, lib=lib),
                                   S.get_power_smoothed(k=k, y=y, dlogk=0.5, idx=idx, param=p), rtol=1e-12)
    for kw in (dict(nmu=9), dict(nmu=5, mu_sampling=False), dict(nmu=7, smooth_dlogk=0.5)):
        a, mu_a = G.power_Kaiser(y=y, kmodes=k, bias=1.7, param=p, lib=lib, **kw)
        b, mu_b = S.power_Kaiser(y=y, kmodes=k, bias=1.7, param=p, **kw)
        np.testing.assert_allclose(mu_a, mu_b, rtol=0, atol=0)
        np.testing.assert_allclose(a, b, rtol=1e-11)
    for N in (512, 257):
        kk = np.geomspace(1e-4, 1e2, N)
        Pk = np.exp(-kk ** 2) + 1e-3 * kk ** -1.5
        for ell in (0, 2, 4):
            xa, ra = G.get_xi_from_P(k=kk, Pk=Pk, ell=ell, lib=lib)
            xb, rb = S.get_xi_from_P(k=kk, Pk=Pk, ell=ell)
            np.testing.assert_allclose(ra, rb, rtol=1e-15)
            assert np.abs(xa - xb).max() <= 1e-10 * np.abs(xb).max(), (N, ell)
    case, pp = _case()
    yy = case["y"][:, -1, :]
    for a, b in zip(G.power_multipoles(y=yy, kmodes=case["kmodes"], b=1.3, param=pp, lib=lib),
                    S.power_multipoles(y=yy, kmodes=case["kmodes"], b=1.3, param=pp)):
        np.testing.assert_allclose(a, b, rtol=1e-13)


def test_product_spectra_source_matches_oracle(emu_lib):
    check_product_spectra(emu_lib)
