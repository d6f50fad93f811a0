"""CPU suite, part 4: the N>1 path with world_size 2 on the gloo backend.  The compute library is
the CPU build of the kernel source (the product library needs a GPU); what is under test is the
host logic: the round-robin deal of k-modes, the padded all-gather and the merge."""
import os
import socket

import numpy as np
import pytest

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_allgather(buf):
    """The host-side gather the package asks its caller for: here torch.distributed on gloo (test plumbing)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(buf))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.numpy() for o in out]


def _worker(rank, world, port, emu_path, q):
    import sys
    sys.path.insert(0, helpers.ROOT)
    sys.path.insert(0, os.path.join(helpers.ROOT, "disco-eb_b200"))
    import torch.distributed as dist
    from discoeb_b200 import _cabi
    from discoeb_b200.distributed import evolve_perturbations_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _cabi.Library(emu_path, prefix="emu_")
    p = helpers.load_tables("fiducial").param()
    y, k, p = evolve_perturbations_sharded(param=p, aexp_out=[0.5, 1.0], kmin=1e-3, kmax=1.0, num_k=11, lib=lib, rank=rank, world=world,
                                           allgather=_gloo_allgather)
    q.put((rank, y, k, p["nout"]))
    dist.barrier()
    dist.destroy_process_group()


def test_package_does_not_import_torch():
    """north_star excludes PyTorch from the product: the package must import and shard without it."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import discoeb_b200.distributed, discoeb_b200.perturbations, discoeb_b200.background; "
            "assert 'torch' not in sys.modules, 'torch imported by the package'" % os.path.join(helpers.ROOT, "disco-eb_b200"))
    subprocess.check_call([sys.executable, "-c", code])


def test_empty_share_and_more_ranks_than_modes(emu_lib):
    """world > num_k leaves ranks without modes: they contribute padding and nobody hangs or fails."""
    from discoeb_b200.distributed import evolve_perturbations_sharded
    p = helpers.load_tables("fiducial").param()
    world = 4
    bufs = {}

    def fake_gather_factory(rank):
        def g(buf):
            return [bufs[r] for r in range(world)]
        return g
    # first pass: collect every rank's buffer by running the local part with a recording gather
    from discoeb_b200 import distributed as D
    for r in range(world):
        rec = {}
        def rec_gather(buf, r=r, rec=rec):
            rec["buf"] = buf.copy()
            raise KeyboardInterrupt          # stop after the local part
        try:
            evolve_perturbations_sharded(param=dict(p), aexp_out=[1.0], kmin=1e-3, kmax=1e-2, num_k=3, lib=emu_lib, rank=r, world=world, allgather=rec_gather)
        except KeyboardInterrupt:
            pass
        bufs[r] = rec["buf"]
    y, k, _ = evolve_perturbations_sharded(param=dict(p), aexp_out=[1.0], kmin=1e-3, kmax=1e-2, num_k=3, lib=emu_lib, rank=3, world=world,
                                           allgather=fake_gather_factory(3))
    ref, _, _ = evolve_perturbations_sharded(param=dict(p), aexp_out=[1.0], kmin=1e-3, kmax=1e-2, num_k=3, lib=emu_lib)
    assert np.array_equal(y, ref)


def test_partition_and_merge_roundtrip():
    from discoeb_b200.distributed import partition_modes, merge_modes
    for nk, w in ((11, 2), (512, 8), (7, 4), (3, 4)):
        parts = partition_modes(nk, w)
        assert sorted(np.concatenate(parts).tolist()) == list(range(nk))
        data = np.arange(nk * 3, dtype=float).reshape(nk, 3)
        per = (nk + w - 1) // w
        padded = []
        for idx in parts:
            b = np.zeros((per, 3))
            b[: len(idx)] = data[idx]
            padded.append(b)
        assert np.array_equal(merge_modes(padded, nk, w), data)


def test_world_size_2_gloo_matches_single_process(emu_lib):
    import torch.multiprocessing as mp
    from discoeb_b200.perturbations import _solve, _kgrid
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib.path, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = helpers.load_tables("fiducial").param()
    ks = _kgrid(1e-3, 1.0, 11, True)
    ref = _solve([p], ks, [0.5, 1.0], lmaxg=11, lmaxgp=11, lmaxr=11, lmaxnu=8, nqmax=3, rtol=1e-4, atol=1e-4, pcoeff=0.25,
                 icoeff=0.8, dcoeff=0.0, factormax=20.0, factormin=0.3, max_steps=2048, return_full=False, device=0, lib=emu_lib)
    for rank, y, k, nout in results:
        assert nout == 2
        np.testing.assert_allclose(k, ks, rtol=1e-15)
        assert np.array_equal(y, ref["y"][0])        # per-mode results do not depend on the sharding


def _worker_jvp(rank, world, port, emu_path, q):
    import sys
    sys.path.insert(0, helpers.ROOT)
    sys.path.insert(0, os.path.join(helpers.ROOT, "disco-eb_b200"))
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    import torch.distributed as dist
    import parity_checks as pc
    from discoeb_b200 import _cabi
    from discoeb_b200.distributed import evolve_perturbations_jvp_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _cabi.Library(emu_path, prefix="emu_")
    case = pc.load_tangent_case("default_n72")
    p = helpers.Tables(case["scalars"], case["tables"], case["nth"], case["nnu"]).param()
    dps = [helpers.Tables(case["d_scalars"][d], case["d_tables"][d], case["nth"], case["nnu"]).param() for d in range(2)]
    y, dy, pk, dpk, k = evolve_perturbations_jvp_sharded(param=p, dparam=dps, aexp_out=[0.5, 1.0], kmin=1e-3, kmax=0.3, num_k=7, lib=lib,
                                                         rank=rank, world=world, allgather=_gloo_allgather)
    q.put((rank, y, dy, pk, dpk, k))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_jvp_matches_single_process(emu_lib):
    """Config-5 path at N>1: (direction, k) items dealt over two ranks == one process."""
    import torch.multiprocessing as mp
    import parity_checks as pc
    from discoeb_b200 import _cabi
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_jvp, args=(r, 2, port, emu_lib.path, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    case = pc.load_tangent_case("default_n72")
    ks = np.geomspace(1e-3, 0.3, 7)
    dims = pc.tangent_dims(case, nk=7, ntan=2, power_idx=4)
    dims.nout = 2
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    ref = emu_lib.evolve_tangent_host(dims, ctrl, case["scalars"][None], case["tables"][None], ks, np.array([0.5, 1.0]),
                                      case["d_scalars"][:2, None], case["d_tables"][:2, None], want_pk=True)
    for rank, y, dy, pk, dpk, k in results:
        np.testing.assert_allclose(k, ks, rtol=1e-15)
        assert np.array_equal(y, ref["y"][0]) and np.array_equal(dy, ref["dy"][:, 0])
        assert np.array_equal(pk, ref["pk"][0]) and np.array_equal(dpk, ref["dpk"][:, 0])
