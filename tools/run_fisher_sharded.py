"""BASELINE config 5 under torchrun (one rank per GPU): P(k) and its Jacobian w.r.t. 7 parameters, 512 k-modes dealt
round-robin over the ranks (every rank integrates all 7 directions of its modes), all-gather at the end.
Rank 0 prints one JSON line per hierarchy size; wall time = max over ranks between barriers."""
import json, os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
import torch, torch.distributed as dist
from discoeb_b200 import _pack
from discoeb_b200.distributed import evolve_perturbations_jvp_sharded

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = np.load(os.path.join(ROOT, "tests", "golden", "fisher_seeds.npz"))

def as_param(scal, tab):
    nth, nnu = int(z["nth"]), int(z["nnu"])
    p = {k: float(scal[i]) for i, k in enumerate(_pack.SCALAR_KEYS)}
    off = 0
    for j, key in enumerate(_pack.SPLINE_KEYS):
        n = nnu if j in (2, 3) else nth
        p[key] = types.SimpleNamespace(x=tab[off:off + n], y=tab[off + n:off + 2 * n], S=tab[off + 2 * n:off + 3 * n])
        off += 3 * n
    return p

param = as_param(z["scalars"], z["tables"])
dparams = [as_param(z["d_scalars"][d], z["d_tables"][d]) for d in range(7)]

def allgather(buf):
    """the gather evolve_perturbations_jvp_sharded asks its caller for: equal-size float64 buffers from every rank, in rank order"""
    t = torch.from_numpy(np.ascontiguousarray(buf)).cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy() for o in outs]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

for dm in ((11, 11, 11, 8, 3), (31, 31, 31, 31, 5)):
    kw = dict(lmaxg=dm[0], lmaxgp=dm[1], lmaxr=dm[2], lmaxnu=dm[3], nqmax=dm[4], max_steps=4096)
    best = 1e30
    for rep in range(3):
        barrier(); t = time.perf_counter()
        y, dy, pk, dpk, k = evolve_perturbations_jvp_sharded(param=dict(param), dparam=dparams, aexp_out=[0.5, 1.0], kmin=1e-4, kmax=10.0,
                                                             num_k=512, device=local, rank=rank, world=world,
                                                             allgather=allgather if world > 1 else None, **kw)
        barrier()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep > 0:
            best = min(best, float(dt[0]))
    if rank == 0:
        n = 7 + dm[0] + dm[1] + dm[2] + 3 + dm[4] * (dm[3] + 1) + 2
        dln = dpk[:, :, -1] / pk[:, -1]
        print(json.dumps(dict(config=f"5: Fisher Jacobian, 7 directions x 512 k, n={n}, a_out=[0.5,1]", n_gpus=world, wall_ms=1e3 * best,
                              items_per_s=7 * 512 / best, finite=bool(np.all(np.isfinite(dy))), checksum=float(np.abs(dln).sum()))), flush=True)
if world > 1:
    dist.destroy_process_group()
