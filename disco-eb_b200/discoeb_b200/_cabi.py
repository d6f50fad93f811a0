"""ctypes binding of the C-ABI declared in ``include/discoeb_b200.h``.

The library is built in-tree by ``__graft_entry__.build()`` (``disco-eb_b200/csrc/Makefile``) as
``disco-eb_b200/discoeb_b200/libdiscoeb_b200.so``.  There is NO fallback: if the shared library
or a CUDA device is missing, loading/calling raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiscoeb_b200.so")


class DebDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "ncosmo", "nk", "nout", "ntan", "lmaxg", "lmaxgp", "lmaxr", "lmaxnu", "nqmax", "nth", "nnu",
        "max_steps", "return_full", "k_per_cosmo", "power_idx", "batch_size")]


class DebCtrl(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "rtol", "atol", "pcoeff", "icoeff", "dcoeff", "factormax", "factormin", "safety")]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class DiscoEBError(RuntimeError):
    pass


class Library:
    """One loaded shared library exposing the ``deb_*`` (or, in the CPU test-suite, ``emu_*``) symbols."""

    def __init__(self, path: str = LIB_PATH, prefix: str = "deb_"):
        if not os.path.exists(path):
            raise DiscoEBError(
                f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
                "discoeb_b200 has no CPU fallback.")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        f = getattr(self.lib, prefix + "evolve_host_f64")
        base = [C.POINTER(DebDims), C.POINTER(DebCtrl), _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _ip]
        f.argtypes = base + ([C.c_int32, C.POINTER(C.c_float)] if prefix == "deb_" else [])
        f.restype = C.c_int
        self._evolve_host = f
        g = getattr(self.lib, prefix + "debug_step_host_f64")
        g.argtypes = [C.POINTER(DebDims), _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp] + ([C.c_int32] if prefix == "deb_" else [])
        g.restype = C.c_int
        self._debug_step = g
        h = getattr(self.lib, prefix + "debug_ics_host_f64")
        h.argtypes = [C.POINTER(DebDims), _dp, _dp, _dp, _dp, _dp, _dp] + ([C.c_int32] if prefix == "deb_" else [])
        h.restype = C.c_int
        self._debug_ics = h
        r = getattr(self.lib, prefix + "debug_replay_host_f64")
        r.argtypes = [C.POINTER(DebDims), C.POINTER(DebCtrl), _dp, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int32, _dp, _ip] + \
            ([C.c_int32] if prefix == "deb_" else [])
        r.restype = C.c_int
        self._debug_replay = r
        tg = getattr(self.lib, prefix + "evolve_tangent_host_f64")
        tg.argtypes = [C.POINTER(DebDims), C.POINTER(DebCtrl)] + [_dp] * 13 + [_ip, _ip, _ip] + \
            ([C.c_int32, C.POINTER(C.c_float)] if prefix == "deb_" else [])
        tg.restype = C.c_int
        self._evolve_tangent_host = tg
        rt = getattr(self.lib, prefix + "debug_replay_tangent_host_f64")
        rt.argtypes = [C.POINTER(DebDims), C.POINTER(DebCtrl)] + [_dp] * 9 + [_ip, _ip, C.c_int32, _dp, _dp, _dp, _ip] + \
            ([C.c_int32] if prefix == "deb_" else [])
        rt.restype = C.c_int
        self._debug_replay_tangent = rt
        bgf = getattr(self.lib, prefix + "background_host_f64")
        bgf.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, C.POINTER(C.c_float)]
        bgf.restype = C.c_int
        self._background_host = bgf
        bgx = getattr(self.lib, prefix + "background_host_ex_f64")
        bgx.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, C.POINTER(C.c_float)]
        bgx.restype = C.c_int
        self._background_host_ex = bgx
        bgl = getattr(self.lib, prefix + "background_extras_len")
        bgl.argtypes = [C.c_int32]
        bgl.restype = C.c_size_t
        self._background_extras_len = bgl
        spf = getattr(self.lib, prefix + "spectra_host_f64")
        spf.argtypes = [C.c_int32, C.c_int32, C.c_int32, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, _dp, C.c_int32, _dp, C.c_int32] + [_dp] * 8
        spf.restype = C.c_int
        self._spectra_host = spf
        if prefix == "deb_":
            self.lib.deb_strerror.restype = C.c_char_p
            self.lib.deb_strerror.argtypes = [C.c_int]
            self.lib.deb_workspace_bytes.restype = C.c_size_t
            self.lib.deb_workspace_bytes.argtypes = [C.POINTER(DebDims)]
            self.lib.deb_table_len.restype = C.c_size_t
            self.lib.deb_table_len.argtypes = [C.POINTER(DebDims)]
            self.lib.deb_nvar.restype = C.c_int32
            self.lib.deb_nvar.argtypes = [C.POINTER(DebDims)]
            self.lib.deb_device_count.restype = C.c_int32
            self.lib.deb_abi_version.restype = C.c_int32
            self.lib.deb_fp64_peak_tflops.restype = C.c_int
            self.lib.deb_fp64_peak_tflops.argtypes = [C.c_int32, _dp, C.POINTER(C.c_float)]
            self.lib.deb_evolve_f64.restype = C.c_int
            self.lib.deb_evolve_f64.argtypes = [C.POINTER(DebDims), C.POINTER(DebCtrl)] + [C.c_void_p] * 10 + \
                [C.c_void_p, C.c_size_t, C.c_void_p]

    # ------------------------------------------------------------------------------------------
    def strerror(self, code: int) -> str:
        if self.prefix == "deb_":
            return self.lib.deb_strerror(code).decode()
        return f"error {code}"

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise DiscoEBError(f"{what} failed: {self.strerror(rc)} (code {rc})")

    @staticmethod
    def nvar(lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax):
        return 7 + (lmaxg + 1) + (lmaxgp + 1) + (lmaxr + 1) + nqmax * (lmaxnu + 1) + 2

    def background_host(self, bg_in, nth: int, device: int = 0, nnu: int = 512, extras: bool = False):
        """Table producer (deb_background_host[_ex]_f64): bg_in[nc, 16] -> (scalars[nc, 24], tables[nc, 3 (5 nth + 2 nnu)], kernel ms)
        and, with ``extras=True``, a fourth item: extras[nc, deb_background_extras_len(nth)] (include/discoeb_b200.h)."""
        bg_in = np.ascontiguousarray(bg_in, dtype=np.float64)
        nc = bg_in.shape[0]
        scal = np.zeros((nc, 24), dtype=np.float64)
        tab = np.zeros((nc, 3 * (5 * nth + 2 * nnu)), dtype=np.float64)
        kms = C.c_float(0.0)
        if not extras:
            self._check(self._background_host(C.c_int32(device), C.c_int32(nc), C.c_int32(nth), _d(bg_in), _d(scal), _d(tab), C.byref(kms)),
                        "background_host_f64")
            return scal, tab, kms.value
        ext = np.zeros((nc, int(self._background_extras_len(C.c_int32(nth)))), dtype=np.float64)
        self._check(self._background_host_ex(C.c_int32(device), C.c_int32(nc), C.c_int32(nth), _d(bg_in), _d(scal), _d(tab), _d(ext),
                                             C.byref(kms)), "background_host_ex_f64")
        return scal, tab, kms.value, ext

    def spectra_host(self, y, k, As, ns, kp, bias, sg_coef, mu, ell, want_xi, device: int = 0):
        """deb_spectra_host_f64 -> dict(P0, P2, P4, Pkmu, Ps_delta, Ps_theta, xi, r) (entries None when not requested)."""
        nk = y.shape[0]
        nmu = 0 if mu is None else len(mu)
        w = 0 if sg_coef is None else len(sg_coef)
        P0, P2, P4 = np.zeros(nk), np.zeros(nk), np.zeros(nk)
        Pkmu = np.zeros((nk, nmu)) if nmu else None
        Psd, Pst = (np.zeros(nk), np.zeros(nk)) if w else (None, None)
        xi, r = (np.zeros(nk), np.zeros(nk)) if want_xi else (None, None)
        self._check(self._spectra_host(C.c_int32(device), C.c_int32(nk), C.c_int32(nmu), _d(y), _d(k), C.c_double(As), C.c_double(ns), C.c_double(kp),
                                       C.c_double(bias), _d(sg_coef), C.c_int32(w), _d(mu), C.c_int32(ell), _d(P0), _d(P2), _d(P4), _d(Pkmu), _d(Psd),
                                       _d(Pst), _d(xi), _d(r)), "spectra_host_f64")
        return dict(P0=P0, P2=P2, P4=P4, Pkmu=Pkmu, Ps_delta=Psd, Ps_theta=Pst, xi=xi, r=r)

    def evolve_sharded_host(self, comm, dims: DebDims, ctrl: DebCtrl, scalars, tables, kmodes, aexp_out, want_pk: bool = False):
        """deb_evolve_sharded_host_f64: the k grid of `dims` dealt over comm.world ranks, NCCL all-gather, full-size results."""
        nc, nk, nout = dims.ncosmo, dims.nk, dims.nout
        scalars = np.ascontiguousarray(scalars, dtype=np.float64); tables = np.ascontiguousarray(tables, dtype=np.float64)
        kmodes = np.ascontiguousarray(kmodes, dtype=np.float64); aexp_out = np.ascontiguousarray(aexp_out, dtype=np.float64)
        y = np.zeros((nc, nk, nout, 20)); pk = np.zeros((nc, nk, nout)) if want_pk else None
        tau_out = np.zeros((nc, nout)); status = np.zeros((nc, nk), dtype=np.int32); nsteps = np.zeros((nc, nk), dtype=np.int32)
        ms = C.c_float(0.0)
        f = self.lib.deb_evolve_sharded_host_f64
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int32, C.POINTER(DebDims), C.POINTER(DebCtrl), _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, C.POINTER(C.c_float)]
        self._check(f(comm.handle, C.c_int32(comm.device), C.byref(dims), C.byref(ctrl), _d(scalars), _d(tables), _d(kmodes), _d(aexp_out),
                      _d(y), _d(pk), _d(tau_out), _i(status), _i(nsteps), C.byref(ms)), "evolve_sharded_host_f64")
        return dict(y=y, pk=pk, tau_out=tau_out, status=status, nsteps=nsteps, elapsed_ms=ms.value)

    def evolve_host(self, dims: DebDims, ctrl: DebCtrl, scalars, tables, kmodes, aexp_out, device: int = 0,
                    want_pk: bool = False):
        """numpy in -> dict of numpy out (y, pk, tau_out, status, nsteps, naccept, kernel_ms)."""
        nc, nk, nout = dims.ncosmo, dims.nk, dims.nout
        nf = self.nvar(dims.lmaxg, dims.lmaxgp, dims.lmaxr, dims.lmaxnu, dims.nqmax) if dims.return_full else 20
        scalars = np.ascontiguousarray(scalars, dtype=np.float64)
        tables = np.ascontiguousarray(tables, dtype=np.float64)
        kmodes = np.ascontiguousarray(kmodes, dtype=np.float64)
        aexp_out = np.ascontiguousarray(aexp_out, dtype=np.float64)
        y = np.zeros((nc, nk, nout, nf), dtype=np.float64)
        pk = np.zeros((nc, nk, nout), dtype=np.float64) if want_pk else None
        tau_out = np.zeros((nc, nout), dtype=np.float64)
        status = np.zeros((nc, nk), dtype=np.int32)
        nsteps = np.zeros((nc, nk), dtype=np.int32)
        nacc = np.zeros((nc, nk), dtype=np.int32)
        args = [C.byref(dims), C.byref(ctrl), _d(scalars), _d(tables), _d(kmodes), _d(aexp_out), _d(y), _d(pk),
                _d(tau_out), _i(status), _i(nsteps), _i(nacc)]
        kms = C.c_float(0.0)
        if self.prefix == "deb_":
            args += [C.c_int32(device), C.byref(kms)]
        self._check(self._evolve_host(*args), "evolve_host_f64")
        return dict(y=y, pk=pk, tau_out=tau_out, status=status, nsteps=nsteps, naccept=nacc, kernel_ms=kms.value)

    def evolve_tangent_host(self, dims: DebDims, ctrl: DebCtrl, scalars, tables, kmodes, aexp_out, d_scalars, d_tables,
                            device: int = 0, want_pk: bool = False, d_kmodes=None):
        """Primal + ``dims.ntan`` forward tangents; d_scalars [ntan, nc, NSCAL], d_tables [ntan, nc, tl],
        optional d_kmodes [ntan] + kmodes.shape."""
        nc, nk, nout, nt = dims.ncosmo, dims.nk, dims.nout, dims.ntan
        nf = self.nvar(dims.lmaxg, dims.lmaxgp, dims.lmaxr, dims.lmaxnu, dims.nqmax) if dims.return_full else 20
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        scalars, tables, kmodes, aexp_out, d_scalars, d_tables = map(f64, (scalars, tables, kmodes, aexp_out, d_scalars, d_tables))
        if d_scalars.shape != (nt,) + scalars.shape or d_tables.shape != (nt,) + tables.shape:
            raise ValueError("tangent seeds must have the primal shapes with a leading [ntan] axis")
        if d_kmodes is not None:
            d_kmodes = f64(d_kmodes)
            if d_kmodes.shape != (nt,) + kmodes.shape:
                raise ValueError("d_kmodes must have the shape of kmodes with a leading [ntan] axis")
        y = np.zeros((nc, nk, nout, nf)); dy = np.zeros((nt, nc, nk, nout, nf))
        pk = np.zeros((nc, nk, nout)) if want_pk else None
        dpk = np.zeros((nt, nc, nk, nout)) if want_pk else None
        tau_out = np.zeros((nc, nout)); dtau_out = np.zeros((nt, nc, nout))
        status = np.zeros((nc, nk), dtype=np.int32); nsteps = np.zeros((nc, nk), dtype=np.int32); nacc = np.zeros((nc, nk), dtype=np.int32)
        args = [C.byref(dims), C.byref(ctrl), _d(scalars), _d(tables), _d(kmodes), _d(aexp_out), _d(d_scalars), _d(d_tables),
                _d(d_kmodes), _d(y), _d(dy), _d(pk), _d(dpk), _d(tau_out), _d(dtau_out), _i(status), _i(nsteps), _i(nacc)]
        kms = C.c_float(0.0)
        if self.prefix == "deb_":
            args += [C.c_int32(device), C.byref(kms)]
        self._check(self._evolve_tangent_host(*args), "evolve_tangent_host_f64")
        return dict(y=y, dy=dy, pk=pk, dpk=dpk, tau_out=tau_out, dtau_out=dtau_out, status=status, nsteps=nsteps, naccept=nacc,
                    kernel_ms=kms.value)

    def debug_replay_tangent(self, dims: DebDims, ctrl: DebCtrl, scalars, tables, kmodes, aexp_out, d_scalars, d_tables,
                             rp_tnext, rp_dtnext, rp_keep, rp_n, device: int = 0, d_kmodes=None):
        nt = dims.ntan
        nf = self.nvar(dims.lmaxg, dims.lmaxgp, dims.lmaxr, dims.lmaxnu, dims.nqmax) if dims.return_full else 20
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        rp_tnext, rp_dtnext = f64(rp_tnext), f64(rp_dtnext)
        rp_keep = np.ascontiguousarray(rp_keep, dtype=np.int32)
        rp_n = np.ascontiguousarray(rp_n, dtype=np.int32)
        y = np.zeros((dims.ncosmo, dims.nk, dims.nout, nf)); dy = np.zeros((nt, dims.ncosmo, dims.nk, dims.nout, nf))
        dtau = np.zeros((nt, dims.ncosmo, dims.nout))
        ns = np.zeros((dims.ncosmo, dims.nk), dtype=np.int32)
        args = [C.byref(dims), C.byref(ctrl), _d(f64(scalars)), _d(f64(tables)), _d(f64(kmodes)), _d(f64(aexp_out)),
                _d(f64(d_scalars)), _d(f64(d_tables)), _d(None if d_kmodes is None else f64(d_kmodes)), _d(rp_tnext), _d(rp_dtnext),
                _i(rp_keep), _i(rp_n),
                C.c_int32(rp_tnext.shape[-1]), _d(y), _d(dy), _d(dtau), _i(ns)]
        if self.prefix == "deb_":
            args.append(C.c_int32(device))
        self._check(self._debug_replay_tangent(*args), "debug_replay_tangent_host_f64")
        return y, dy, dtau, ns

    def debug_step(self, dims: DebDims, scalars, tables, kmodes, t0, t1, y0, device: int = 0):
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        y1 = np.zeros_like(y0)
        err = np.zeros_like(y0)
        t0 = np.ascontiguousarray(t0, dtype=np.float64)
        t1 = np.ascontiguousarray(t1, dtype=np.float64)
        args = [C.byref(dims), _d(np.ascontiguousarray(scalars)), _d(np.ascontiguousarray(tables)),
                _d(np.ascontiguousarray(kmodes, dtype=np.float64)), _d(t0), _d(t1), _d(y0), _d(y1), _d(err)]
        if self.prefix == "deb_":
            args.append(C.c_int32(device))
        self._check(self._debug_step(*args), "debug_step_host_f64")
        return y1, err

    def debug_ics(self, dims: DebDims, scalars, tables, kmodes, aexp_out, device: int = 0):
        n = self.nvar(dims.lmaxg, dims.lmaxgp, dims.lmaxr, dims.lmaxnu, dims.nqmax)
        ts = np.zeros((dims.ncosmo, dims.nk), dtype=np.float64)
        y0 = np.zeros((dims.ncosmo, dims.nk, n), dtype=np.float64)
        args = [C.byref(dims), _d(np.ascontiguousarray(scalars)), _d(np.ascontiguousarray(tables)),
                _d(np.ascontiguousarray(kmodes, dtype=np.float64)), _d(np.ascontiguousarray(aexp_out, dtype=np.float64)),
                _d(ts), _d(y0)]
        if self.prefix == "deb_":
            args.append(C.c_int32(device))
        self._check(self._debug_ics(*args), "debug_ics_host_f64")
        return ts, y0

    def debug_replay(self, dims: DebDims, ctrl: DebCtrl, scalars, tables, kmodes, aexp_out, rp_tnext, rp_keep, rp_n,
                     device: int = 0):
        """Integrate along a prescribed step sequence (see include/discoeb_b200.h)."""
        nf = self.nvar(dims.lmaxg, dims.lmaxgp, dims.lmaxr, dims.lmaxnu, dims.nqmax) if dims.return_full else 20
        rp_tnext = np.ascontiguousarray(rp_tnext, dtype=np.float64)
        rp_keep = np.ascontiguousarray(rp_keep, dtype=np.int32)
        rp_n = np.ascontiguousarray(rp_n, dtype=np.int32)
        y = np.zeros((dims.ncosmo, dims.nk, dims.nout, nf), dtype=np.float64)
        ns = np.zeros((dims.ncosmo, dims.nk), dtype=np.int32)
        args = [C.byref(dims), C.byref(ctrl), _d(np.ascontiguousarray(scalars)), _d(np.ascontiguousarray(tables)),
                _d(np.ascontiguousarray(kmodes, dtype=np.float64)), _d(np.ascontiguousarray(aexp_out, dtype=np.float64)),
                _d(rp_tnext), _i(rp_keep), _i(rp_n), C.c_int32(rp_tnext.shape[-1]), _d(y), _i(ns)]
        if self.prefix == "deb_":
            args.append(C.c_int32(device))
        self._check(self._debug_replay(*args), "debug_replay_host_f64")
        return y, ns

    def fp64_peak_tflops(self, device: int = 0):
        t = C.c_double(0.0)
        mhz = C.c_float(0.0)
        self._check(self.lib.deb_fp64_peak_tflops(device, C.byref(t), C.byref(mhz)), "fp64_peak_tflops")
        return t.value, mhz.value


_default = None


def default_library() -> Library:
    global _default
    if _default is None:
        _default = Library()
    return _default


def make_dims(*, ncosmo, nk, nout, lmaxg, lmaxgp, lmaxr, lmaxnu, nqmax, nth, nnu, max_steps, return_full=False,
              k_per_cosmo=False, power_idx=-1, ntan=0, batch_size=0) -> DebDims:
    return DebDims(ncosmo=ncosmo, nk=nk, nout=nout, ntan=ntan, lmaxg=lmaxg, lmaxgp=lmaxgp, lmaxr=lmaxr, lmaxnu=lmaxnu,
                   nqmax=nqmax, nth=nth, nnu=nnu, max_steps=max_steps, return_full=int(bool(return_full)),
                   k_per_cosmo=int(bool(k_per_cosmo)), power_idx=power_idx, batch_size=batch_size)


def make_ctrl(*, rtol, atol, pcoeff=0.25, icoeff=0.8, dcoeff=0.0, factormax=20.0, factormin=0.3, safety=0.9) -> DebCtrl:
    return DebCtrl(rtol=rtol, atol=atol, pcoeff=pcoeff, icoeff=icoeff, dcoeff=dcoeff, factormax=factormax,
                   factormin=factormin, safety=safety)
