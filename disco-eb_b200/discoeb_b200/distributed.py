"""Multi-GPU driver: one process per GPU, k-modes dealt round-robin, one final all-gather.

The reference has no multi-device code at all; every (cosmology, k) mode
is an independent ODE solve, so the path shards without any data-path collective.  Modes are
dealt round-robin over the ascending k grid (mode i -> rank i mod world) because the cost of a
mode grows steeply with k; the only communication is the final gather of the
``[nk_local, nout, 20]`` transfer functions (80 KB per rank for 512 modes): ``all_gather`` over
NCCL/NVLink on GPUs, gloo in the CPU test-suite.

``torch.distributed`` is plumbing only; the solve goes through the C-ABI like the single-GPU path.
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


def evolve_perturbations_sharded(*, param, aexp_out, kmin, kmax, num_k, group=None, lib=None, device=None, **kw):
    """``evolve_perturbations`` with the k grid sharded over the ranks of ``group``.

    Every rank returns the full ``(y[num_k, nout, 20|n], kmodes, param)``.  Must be called by all
    ranks of the group (it ends in a collective).  ``lib`` injects the compute library (the CPU
    tests pass the emulation build); by default the CUDA library on ``device`` (= local rank).
    """
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dologk = kw.pop("dologk", True)
    throw = kw.pop("throw", True)
    kmodes = _pt._kgrid(kmin, kmax, num_k, dologk)
    mine = partition_modes(num_k, world)[rank]
    per = (num_k + world - 1) // world
    args = dict(lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, rtol=1e-4, atol=1e-4, pcoeff=0.25, icoeff=0.80,
                dcoeff=0.0, factormax=20.0, factormin=0.3, max_steps=2048, return_full=False)
    args.update(kw)
    dev = device if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    out = _pt._solve([param], kmodes[mine], aexp_out, device=dev, lib=lib, **args)
    y = out["y"][0]
    nf, nout = y.shape[-1], y.shape[1]
    # pad to a common length so that one all_gather moves everything (status rides along as a field)
    buf = np.zeros((per, nout, nf + 1))
    buf[: len(mine), :, :nf] = y
    buf[: len(mine), 0, nf] = out["status"][0]
    use_cuda = dist.is_initialized() and dist.get_backend(group) == "nccl"
    t = torch.from_numpy(buf)
    if use_cuda:
        t = t.cuda(dev)
    if world > 1:
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        parts = [g.cpu().numpy() for g in gathered]
    else:
        parts = [t.cpu().numpy()]
    full = merge_modes(parts, num_k, world)
    status = full[:, 0, nf].astype(np.int32)
    _pt._check_status(status, None, args["max_steps"], throw)
    param["lmaxg"], param["lmaxgp"], param["lmaxr"] = args["lmaxg"], args["lmaxgp"], args["lmaxr"]
    param["lmaxnu"], param["nqmax"] = args["lmaxnu"], args["nqmax"]
    param["nout"], param["tau_out"] = nout, out["tau_out"][0]
    return full[:, :, :nf], kmodes, param


def evolve_perturbations_jvp_sharded(*, param, dparam, aexp_out, kmin, kmax, num_k, group=None, lib=None, device=None,
                                     power_idx: int = 4, **kw):
    """``evolve_perturbations_jvp`` (the Fisher / ``jacfwd`` workload, BASELINE config 5) with the k grid dealt
    round-robin over the ranks of ``group``: every rank integrates all directions of its own modes -- (direction, k)
    work items are independent -- and one padded all-gather returns everything to every rank.

    Returns ``(y[num_k, nout, 20], dy[ntan, num_k, nout, 20], pk[num_k, nout], dpk[ntan, num_k, nout], kmodes)``.
    """
    import torch
    import torch.distributed as dist
    from . import _cabi
    from ._pack import pack_params, pack_tangent
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
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
    dev = device if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    dims = _cabi.make_dims(ncosmo=1, nk=len(mine), nout=nout, lmaxg=args["lmaxg"], lmaxgp=args["lmaxgp"], lmaxr=args["lmaxr"],
                           lmaxnu=args["lmaxnu"], nqmax=args["nqmax"], nth=nth, nnu=nnu, max_steps=args["max_steps"],
                           power_idx=power_idx, ntan=nt)
    ctrl = _cabi.make_ctrl(rtol=args["rtol"], atol=args["atol"], pcoeff=args["pcoeff"], icoeff=args["icoeff"], dcoeff=args["dcoeff"],
                           factormax=args["factormax"], factormin=args["factormin"])
    out = lib.evolve_tangent_host(dims, ctrl, scalars, tables, kmodes[mine], aexp_out, d_scalars, d_tables, device=dev, want_pk=True)
    # one buffer per rank: [per, nout, (1 + nt) * 21 + 1] = y | pk | dy_d | dpk_d ... | status
    w = (1 + nt) * 21 + 1
    buf = np.zeros((per, nout, w))
    m = len(mine)
    buf[:m, :, :20] = out["y"][0]
    buf[:m, :, 20] = out["pk"][0]
    for d in range(nt):
        o = 21 * (1 + d)
        buf[:m, :, o:o + 20] = out["dy"][d, 0]
        buf[:m, :, o + 20] = out["dpk"][d, 0]
    buf[:m, 0, w - 1] = out["status"][0]
    use_cuda = dist.is_initialized() and dist.get_backend(group) == "nccl"
    t = torch.from_numpy(buf)
    if use_cuda:
        t = t.cuda(dev)
    if world > 1:
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        parts = [g.cpu().numpy() for g in gathered]
    else:
        parts = [t.cpu().numpy()]
    full = merge_modes(parts, num_k, world)
    _pt._check_status(full[:, 0, w - 1].astype(np.int32), None, args["max_steps"], throw)
    y, pk = full[:, :, :20], full[:, :, 20]
    dy = np.stack([full[:, :, 21 * (1 + d):21 * (1 + d) + 20] for d in range(nt)])
    dpk = np.stack([full[:, :, 21 * (1 + d) + 20] for d in range(nt)])
    param["nout"], param["tau_out"] = nout, out["tau_out"][0]
    return y, dy, pk, dpk, kmodes
