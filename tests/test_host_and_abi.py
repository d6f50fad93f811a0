"""CPU suite, part 3: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from discoeb_b200 import _cabi, _pack

ROOT = helpers.ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "discoeb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(deb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "build the CUDA library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/discoeb_b200.h but not exported"


def test_library_metadata_calls_without_gpu():
    lib = _cabi.Library()
    assert lib.lib.deb_abi_version() == 2
    d = _cabi.make_dims(ncosmo=1, nk=4, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=256, nnu=512, max_steps=10)
    assert lib.lib.deb_nvar(ctypes.byref(d)) == 265
    assert lib.lib.deb_table_len(ctypes.byref(d)) == 3 * (5 * 256 + 2 * 512)
    assert lib.lib.deb_workspace_bytes(ctypes.byref(d)) >= 4
    assert lib.strerror(-2).startswith("unsupported")


def test_product_path_fails_loudly_without_gpu(tables):
    """No CPU fallback: without a device the host entry returns DEB_E_NODEVICE and the Python API raises."""
    lib = _cabi.Library()
    if lib.lib.deb_device_count() > 0:
        pytest.skip("a GPU is visible")
    from discoeb_b200.perturbations import evolve_perturbations
    with pytest.raises(_cabi.DiscoEBError, match="no CUDA device"):
        evolve_perturbations(param=tables["fiducial"].param(), aexp_out=[1.0], kmin=1e-3, kmax=1.0, num_k=4)


def test_missing_library_raises(tmp_path):
    with pytest.raises(_cabi.DiscoEBError, match="no CPU fallback"):
        _cabi.Library(str(tmp_path / "nope.so"))


def test_pack_roundtrip_and_validation(tables):
    tab = tables["fiducial"]
    p = tab.param()
    scal, tb, nth, nnu = _pack.pack_param(p)
    assert np.array_equal(scal, tab.scalars) and np.array_equal(tb, tab.tables) and (nth, nnu) == (tab.nth, tab.nnu)

    class RefSpline:                      # attribute names of the reference's pytree (spline_interpolation.py:111-113)
        def __init__(self, s):
            self._x_, self._y_, self._S_full_ = s.x, s.y, s.S
    q = dict(p)
    for key in _pack.SPLINE_KEYS:
        q[key] = RefSpline(p[key])
    scal2, tb2, _, _ = _pack.pack_param(q)
    assert np.array_equal(tb2, tb)
    bad = dict(p)
    del bad["grhom"]
    with pytest.raises(KeyError):
        _pack.pack_param(bad)
    bad = dict(p)
    bad["xe_of_loga_spline"] = p["a_of_tau_spline"]
    with pytest.raises(ValueError):
        _pack.pack_param(bad)


def test_argument_validation_needs_no_gpu(emu_lib, tables):
    tab = tables["fiducial"]
    ks = np.array([0.1])
    for kw in (dict(lmaxg=2), dict(nqmax=6), dict(nqmax=2), dict(lmaxnu=200)):
        args = dict(ncosmo=1, nk=1, nout=1, lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, nth=tab.nth, nnu=tab.nnu, max_steps=10)
        args.update(kw)
        with pytest.raises(_cabi.DiscoEBError):
            emu_lib.evolve_host(_cabi.make_dims(**args), _cabi.make_ctrl(rtol=1e-4, atol=1e-4), tab.scalars[None], tab.tables[None], ks, np.array([1.0]))


def test_unsorted_outputs_rejected(tables):
    from discoeb_b200.perturbations import evolve_perturbations
    with pytest.raises(ValueError, match="ascending"):
        evolve_perturbations(param=tables["fiducial"].param(), aexp_out=[1.0, 0.5], kmin=1e-3, kmax=1.0, num_k=4)


def test_host_api_signatures_match_reference():
    """Drop-in check: every keyword of the reference's entry points (perturbations.py:926-1224, recorded by
    tools/make_reference_signatures.py) exists here in the same order with the same default; the mirror may
    only append keywords of its own (device, throw, ...)."""
    import inspect
    import json
    from discoeb_b200 import perturbations

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_signatures.json")))
    assert len(ref) == 7
    for name, sig in ref.items():
        fn = getattr(perturbations, name)  # as in the reference: all seven live in `perturbations`
        p = inspect.signature(fn).parameters
        assert [q.name for q in p.values() if q.kind is q.POSITIONAL_OR_KEYWORD] == sig["positional"], name
        mine = [(q.name, q.default) for q in p.values() if q.kind is q.KEYWORD_ONLY]
        want = sig["keyword_only"]
        assert [m[0] for m in mine[:len(want)]] == [w[0] for w in want], name
        for (kw, default), (_, have) in zip(want, mine):
            if default is None:
                # required in the reference; get_xi_from_P's N may default to None here
                assert have is inspect.Parameter.empty or have is None, (name, kw)
            else:
                assert have == eval(default), (name, kw, have, default)
        # additions must be optional
        assert all(d is not inspect.Parameter.empty for _, d in mine[len(want):]), name
