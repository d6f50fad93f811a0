"""Multi-GPU driver: one process per GPU, k-modes dealt round-robin, one final all-gather.

The reference has no multi-device code at all; every (cosmology, k) mode
is an independent ODE solve, so the path shards without any data-path collective.  Modes are
dealt round-robin over the ascending k grid (mode i -> rank i mod world) because the cost of a
mode grows steeply with k; the only communication is the final gather of the
``[nk_local, nout, 20]`` transfer functions (80 KB per rank for 512 modes).

Two ways to run it, neither of which makes this package import a tensor framework:

* ``comm=NativeComm(...)``: the library's own multi-GPU entry ``deb_evolve_sharded_f64`` (csrc/deb_dist.cu): NCCL
  all-gather on the device (or, with device pointers, the peer-store epilogue), communicator created by the library
  from an id the caller broadcasts with whatever it has (``broadcast`` callable: MPI, torch.distributed, a file);
* ``allgather=callable``: host-side gather supplied by the caller (``allgather(ndarray) -> [ndarray per rank]``) --
  what the CPU test-suite uses with the gloo backend and the CPU build of the kernel source.
"""
from __future__ import annotations

import numpy as np

from . import perturbations as _pt


def partition_modes(num_k: int, world: int):
    """Index sets of the round-robin deal: rank r owns modes r, r+world, r+2 world, ..."""
    return [np.arange(r, num_k, world) for r in range(world)]


def merge_modes(parts, num_k: int, world: int):
    """Inverse of :func:`partition_modes` for per-rank result arrays (mode axis first)."""
    first = np.asarray(parts[0])
    out = np.empty((num_k,) + first.shape[1:], dtype=first.dtype)
    for r, idx in enumerate(partition_modes(num_k, world)):
        out[idx] = np.asarray(parts[r])[: len(idx)]
    return out




class NativeComm:
    """The library's NCCL communicator for ``world`` ranks of one box (one process per GPU, current device = ``device``).
    ``broadcast(payload_or_None) -> payload`` must return rank 0's 128-byte id on every rank."""

    def __init__(self, world: int, rank: int, broadcast, device: int = 0, lib=None):
        import ctypes as C
        from . import _cabi
        self.lib = lib or _cabi.default_library()
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        buf = C.create_string_buffer(128)
        if rank == 0:
            self.lib._check(self.lib.lib.deb_nccl_unique_id(buf), "nccl_unique_id")
        ident = broadcast(bytes(buf.raw) if rank == 0 else None)
        self.handle = C.c_void_p()
        self.lib._check(self.lib.lib.deb_comm_create_on(C.c_int32(device), C.c_int32(world), C.c_int32(rank), C.c_char_p(ident),
                                                       C.byref(self.handle)), "comm_create")

    def close(self):
        if self.handle:
            self.lib.lib.deb_comm_destroy(self.handle)
            self.handle = None


def _gather_rows(buf, rank, world, allgather):
    if world == 1:
        return [buf]
    if allgather is None:
        raise ValueError("world > 1 needs comm=NativeComm(...) or an allgather callable")
    return [np.asarray(p) for p in allgather(buf)]


def evolve_perturbations_sharded(*, param, aexp_out, kmin, kmax, num_k, rank: int = 0, world: int = 1, comm: NativeComm | None = None,
                                 allgather=None, lib=None, device=None, **kw):
    """``evolve_perturbations`` with the k grid dealt round-robin over ``world`` ranks.

    Every rank returns the full ``(y[num_k, nout, 20], kmodes, param)`` and must make the call (it ends in a
    collective).  A rank whose share is empty, or whose local solve raised, still takes part in the gather; errors
    travel with the rows and are raised on every rank afterwards.
    """
    from . import _cabi
    dologk = kw.pop("dologk", True)
    throw = kw.pop("throw", True)
    kmodes = _pt._kgrid(kmin, kmax, num_k, dologk)
    args = dict(lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, rtol=1e-4, atol=1e-4, pcoeff=0.25, icoeff=0.80,
                dcoeff=0.0, factormax=20.0, factormin=0.3, max_steps=2048, return_full=False)
    args.update(kw)
    if args["return_full"]:
        raise NotImplementedError("the sharded driver returns the 20 output fields")
    aexp = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    nout = aexp.size
    dev = device if device is not None else (comm.device if comm is not None else 0)
    if comm is not None:
        from ._pack import pack_params
        scalars, tables, nth, nnu = pack_params([param])
        dims = _cabi.make_dims(ncosmo=1, nk=num_k, nout=nout, lmaxg=args["lmaxg"], lmaxgp=args["lmaxgp"], lmaxr=args["lmaxr"],
                               lmaxnu=args["lmaxnu"], nqmax=args["nqmax"], nth=nth, nnu=nnu, max_steps=args["max_steps"], power_idx=-1)
        ctrl = _cabi.make_ctrl(rtol=args["rtol"], atol=args["atol"], pcoeff=args["pcoeff"], icoeff=args["icoeff"], dcoeff=args["dcoeff"],
                               factormax=args["factormax"], factormin=args["factormin"])
        out = comm.lib.evolve_sharded_host(comm, dims, ctrl, scalars, tables, kmodes, aexp)
        y, status, tau_out = out["y"][0], out["status"][0], out["tau_out"][0]
    else:
        mine = partition_modes(num_k, world)[rank]
        per = (num_k + world - 1) // world
        buf = np.zeros((per, nout, 22))           # 20 fields | status | error flag of the rank
        tau_out = None
        try:
            if len(mine):
                out = _pt._solve([param], kmodes[mine], aexp, device=dev, lib=lib, **args)
                buf[: len(mine), :, :20] = out["y"][0]
                buf[: len(mine), 0, 20] = out["status"][0]
                tau_out = out["tau_out"][0]
        except Exception:          # noqa: BLE001 -- reported after the collective so that no rank is left hanging in it
            buf[:, :, 21] = 1.0
        parts = _gather_rows(buf, rank, world, allgather)
        if any(np.any(p[..., 21] != 0) for p in parts):
            raise RuntimeError("a rank failed in its local solve (see that rank's log)")
        full = merge_modes(parts, num_k, world)
        y, status = full[:, :, :20], full[:, 0, 20].astype(np.int32)
        if tau_out is None:
            tau_out = np.asarray(param["tau_of_a_spline"].evaluate(aexp))
    _pt._check_status(status, None, args["max_steps"], throw)
    param["lmaxg"], param["lmaxgp"], param["lmaxr"] = args["lmaxg"], args["lmaxgp"], args["lmaxr"]
    param["lmaxnu"], param["nqmax"] = args["lmaxnu"], args["nqmax"]
    param["nout"], param["tau_out"] = nout, tau_out
    return y, kmodes, param


def evolve_perturbations_jvp_sharded(*, param, dparam, aexp_out, kmin, kmax, num_k, rank: int = 0, world: int = 1, allgather=None,
                                     lib=None, device: int = 0, power_idx: int = 4, **kw):
    """``evolve_perturbations_jvp`` (the Fisher / ``jacfwd`` workload, BASELINE config 5) with the k grid dealt
    round-robin over ``world`` ranks: every rank integrates all directions of its own modes -- (direction, k) work
    items are independent -- and one padded all-gather (``allgather`` callable, see the module docstring) returns
    everything to every rank.

    Returns ``(y[num_k, nout, 20], dy[ntan, num_k, nout, 20], pk[num_k, nout], dpk[ntan, num_k, nout], kmodes)``.
    """
    from . import _cabi
    from ._pack import pack_params, pack_tangent
    dologk = kw.pop("dologk", True)
    throw = kw.pop("throw", True)
    args = dict(lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, rtol=1e-4, atol=1e-4, pcoeff=0.25, icoeff=0.80, dcoeff=0.0,
                factormax=20.0, factormin=0.3, max_steps=2048)
    args.update(kw)
    lib = lib or _cabi.default_library()
    dlist = [dparam] if isinstance(dparam, dict) else list(dparam)
    nt = len(dlist)
    kmodes = _pt._kgrid(kmin, kmax, num_k, dologk)
    mine = partition_modes(num_k, world)[rank]
    per = (num_k + world - 1) // world
    scalars, tables, nth, nnu = pack_params([param])
    seeds = [pack_tangent(param, d) for d in dlist]
    d_scalars = np.stack([s_[0] for s_ in seeds])[:, None]
    d_tables = np.stack([s_[1] for s_ in seeds])[:, None]
    aexp_out = np.atleast_1d(np.asarray(aexp_out, dtype=np.float64))
    nout = aexp_out.size
    # one buffer per rank: [per, nout, (1 + nt) * 21 + 2] = y | pk | dy_d | dpk_d ... | status | error flag
    w = (1 + nt) * 21 + 2
    buf = np.zeros((per, nout, w))
    m = len(mine)
    tau_out = None
    try:
        if m:
            dims = _cabi.make_dims(ncosmo=1, nk=m, nout=nout, lmaxg=args["lmaxg"], lmaxgp=args["lmaxgp"], lmaxr=args["lmaxr"],
                                   lmaxnu=args["lmaxnu"], nqmax=args["nqmax"], nth=nth, nnu=nnu, max_steps=args["max_steps"],
                                   power_idx=power_idx, ntan=nt)
            ctrl = _cabi.make_ctrl(rtol=args["rtol"], atol=args["atol"], pcoeff=args["pcoeff"], icoeff=args["icoeff"], dcoeff=args["dcoeff"],
                                   factormax=args["factormax"], factormin=args["factormin"])
            out = lib.evolve_tangent_host(dims, ctrl, scalars, tables, kmodes[mine], aexp_out, d_scalars, d_tables, device=device, want_pk=True)
            buf[:m, :, :20] = out["y"][0]
            buf[:m, :, 20] = out["pk"][0]
            for d in range(nt):
                o = 21 * (1 + d)
                buf[:m, :, o:o + 20] = out["dy"][d, 0]
                buf[:m, :, o + 20] = out["dpk"][d, 0]
            buf[:m, 0, w - 2] = out["status"][0]
            tau_out = out["tau_out"][0]
    except Exception:              # noqa: BLE001 -- reported after the collective
        buf[:, :, w - 1] = 1.0
    parts = _gather_rows(buf, rank, world, allgather)
    if any(np.any(p[..., w - 1] != 0) for p in parts):
        raise RuntimeError("a rank failed in its local solve (see that rank's log)")
    full = merge_modes(parts, num_k, world)
    _pt._check_status(full[:, 0, w - 2].astype(np.int32), None, args["max_steps"], throw)
    y, pk = full[:, :, :20], full[:, :, 20]
    dy = np.stack([full[:, :, 21 * (1 + d):21 * (1 + d) + 20] for d in range(nt)])
    dpk = np.stack([full[:, :, 21 * (1 + d) + 20] for d in range(nt)])
    param["nout"] = nout
    param["tau_out"] = tau_out if tau_out is not None else np.asarray(param["tau_of_a_spline"].evaluate(aexp_out))
    return y, dy, pk, dpk, kmodes
