"""Multi-GPU (one box) run of the library's sharded entry deb_evolve_sharded_f64 under torchrun: BASELINE config 3
(w0wa + massive nu, n = 265, 4096 k dealt round-robin) and config 5-like small grids, with BOTH ways of assembling the
result on every rank -- gather=0 (ncclAllGather, the library's own communicator) and gather=1 (peer-store epilogue over
NVLink, buffers from torch symmetric memory).  torch is the harness here (device buffers, rendezvous), not the product.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_sharded.py [nk]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from discoeb_b200 import _cabi  # noqa: E402
from discoeb_b200.distributed import NativeComm  # noqa: E402


def main():
    nk = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.default_library()
    L = lib.lib

    def bcast(payload):
        box = [payload]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    comm = NativeComm(world, rank, bcast, device=local, lib=lib)
    tab = helpers.load_tables("w0wa")
    ks = np.geomspace(1e-4, 10.0, nk)
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=1, lmaxg=31, lmaxgp=31, lmaxr=31, lmaxnu=31, nqmax=5, nth=tab.nth, nnu=tab.nnu, max_steps=4096, power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=1e-4, atol=1e-4)
    f64 = dict(dtype=torch.float64, device=dev)
    d_sc = torch.from_numpy(tab.scalars[None].copy()).to(dev); d_tb = torch.from_numpy(tab.tables[None].copy()).to(dev)
    d_k = torch.from_numpy(ks).to(dev); d_a = torch.tensor([1.0], **f64); d_tau = torch.zeros(1, **f64)
    L.deb_sharded_workspace_bytes.restype = C.c_size_t
    L.deb_sharded_workspace_bytes.argtypes = [C.POINTER(_cabi.DebDims), C.c_int32]
    wsb = L.deb_sharded_workspace_bytes(C.byref(dims), world)
    d_ws = torch.zeros(wsb // 8 + 8, **f64)
    L.deb_evolve_sharded_f64.restype = C.c_int
    L.deb_evolve_sharded_f64.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(_cabi.DebDims), C.POINTER(_cabi.DebCtrl)] + [C.c_void_p] * 9 + \
        [C.c_void_p, C.c_size_t, C.c_int32] + [C.c_void_p] * 4 + [C.c_void_p]
    results = {}
    # result buffers: symmetric memory so that every rank can address every other rank's copy
    try:
        import torch.distributed._symmetric_memory as symm
        nrow = nk
        sy = symm.empty(nrow * 20, **f64); spk = symm.empty(nrow, **f64)
        sst = symm.empty(nrow, dtype=torch.int32, device=dev); sns = symm.empty(nrow, dtype=torch.int32, device=dev)
        hy, hpk, hst, hns = (symm.rendezvous(t, dist.group.WORLD) for t in (sy, spk, sst, sns))
        peers = [(C.c_void_p * world)(*[int(p) for p in h.buffer_ptrs]) for h in (hy, hpk, hst, hns)]
        have_symm = True
    except Exception as e:          # noqa: BLE001
        if rank == 0:
            print("symmetric memory unavailable:", repr(e)[:200], flush=True)
        have_symm = False
    y0 = torch.zeros(nk * 20, **f64); pk0 = torch.zeros(nk, **f64); st0 = torch.zeros(nk, dtype=torch.int32, device=dev); ns0 = torch.zeros(nk, dtype=torch.int32, device=dev)

    def run(gather, y, pk, stt, nst, reps=4):
        best = 1e9
        for _ in range(reps):
            y.zero_(); stt.fill_(7)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p = peers if gather == 1 else [None] * 4
            rc = L.deb_evolve_sharded_f64(comm.handle, world, rank, C.byref(dims), C.byref(ctrl), d_sc.data_ptr(), d_tb.data_ptr(), d_k.data_ptr(),
                                          d_a.data_ptr(), y.data_ptr(), pk.data_ptr(), d_tau.data_ptr(), stt.data_ptr(), nst.data_ptr(),
                                          d_ws.data_ptr(), C.c_size_t(wsb), gather, p[0], p[1], p[2], p[3], C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, lib.strerror(rc)
            if gather == 1:
                torch.cuda.synchronize(); hy.barrier()          # peer stores are complete once every rank's stream is done
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], **f64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        return best
    ms0 = run(0, y0, pk0, st0, ns0)
    assert int(st0.abs().max()) == 0, "modes failed / missing after the all-gather"
    results["allgather_ms"] = ms0
    if have_symm:
        ms1 = run(1, sy, spk, sst, sns)
        assert int(sst.abs().max()) == 0, "modes failed / missing after the peer stores"
        same = bool(torch.equal(sy, y0) and torch.equal(spk, pk0) and torch.equal(sns, ns0))
        results["peer_store_ms"] = ms1
        results["peer_equals_allgather"] = same
    chk = torch.tensor([float(pk0.sum())], **f64)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    results["every_rank_same_checksum"] = bool(all(float(c) == float(allc[0]) for c in allc))
    if rank == 0:
        out = dict(what=f"config 3 sharded: w0wa n=265, {nk} k dealt round-robin", n_gpus=world, modes=nk, steps=int(ns0.sum()), **results,
                   modes_per_s_allgather=nk / ms0 * 1e3, modes_per_s_peer=(nk / results["peer_store_ms"] * 1e3 if have_symm else None))
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"sharded_{world}gpu_{nk}.json"), "w") as f:
            json.dump(out, f)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
