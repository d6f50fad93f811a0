#!/usr/bin/env python
"""bench.py -- headline benchmark of the Einstein-Boltzmann hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic input: every k-mode of
BASELINE.json configs[1] (LambdaCDM + one massive neutrino, lmax=31 -> 32 multipoles per
hierarchy, nq=5, n=265 equations per mode, num_k=512, k in [1e-4, 10]/Mpc, a_out=1,
rtol=atol=1e-4, reference solver defaults) integrated from its start time to z=0 and converted
to the 20 output fields + P(k).  Inputs are the committed fiducial tables (tests/golden).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU): every rank integrates its own 512-mode batch
(weak scaling: independent cosmologies, no data-path collective) and the transfer functions are
all-gathered over NCCL at the end of each step, inside the timed region.

`value`   : k-modes/s, whole job, inputs resident in HBM, CUDA-event time of the step (max over ranks).
`e2e`     : same metric through the host C-ABI entry (deb_evolve_host_f64) with HOST buffers:
            H2D of scalars/tables/k, kernel, D2H of y/P(k)/status inside the timed region.
`roofline`: FP64-FMA pipe (this path is neither HBM- nor tensor-bound): algorithmic flops
            F_step(n) = 370 n + 3000 per attempted step x attempted steps,
            over the kernel's CUDA-event time, against the DFMA peak measured on this GPU by
            deb_fp64_peak_tflops (MEASURED_PEAKS.json holds no FP64 figure).
`cpu_baseline`: the restated reference (NumPy/LAPACK oracle, dense Jacobian + dense LU, lock-step
            over modes like the reference's vmap) on the host cores, on a bounded sub-sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disco-eb_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="config2: LCDM + 1 massive nu, lmax=31 (32 multipoles), nq=5, n=265, num_k=512, "
                     "k in [1e-4,10]/Mpc, a_out=1, rtol=atol=1e-4",
                dims=(31, 31, 31, 31, 5), nk=512, kmin=1e-4, kmax=10.0, aexp_out=[1.0], rtol=1e-4, max_steps=2048)
METRIC = "k-modes/sec (ms per P(k), N_k=512, in ms_per_step)"


# dram__bytes_read.sum + dram__bytes_write.sum of one k_evolve_team<3,4,2> launch on this workload (ncu --set full),
# read from the committed summary so that the number and its evidence cannot drift apart
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r1_v24_k_evolve_team_ncu_summary.txt")


def ncu_dram_bytes_per_launch():
    try:
        tot = 0.0
        for ln in open(NCU_SUMMARY):
            for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
                if ln.startswith(key):
                    unit = ln[len(key):ln.index("]")]
                    val = float(ln.split("=")[1])
                    tot += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return int(tot) if tot > 0 else None
    except Exception:
        return None


def f_step(n):
    return 370.0 * n + 3000.0


def load_tables():
    import helpers
    return helpers.load_tables("fiducial")


# ------------------------------------------------------------------------------------------------
# CPU arm: the restated reference on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(tab, stride):
    """One bounded pass: every `stride`-th mode of the workload through the NumPy oracle."""
    import oracle.discoeb_oracle as O
    p = tab.param()
    ks = np.geomspace(WORKLOAD["kmin"], WORKLOAD["kmax"], WORKLOAD["nk"])[stride // 2::stride]
    lg, lp, lr, ln, nq = WORKLOAD["dims"]
    t = time.perf_counter()
    y, k, _, info = O.evolve_perturbations(param=p, aexp_out=WORKLOAD["aexp_out"], kmin=0, kmax=0, num_k=len(ks), kmodes=ks,
                                           lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, rtol=WORKLOAD["rtol"],
                                           atol=WORKLOAD["rtol"], max_steps=WORKLOAD["max_steps"], return_info=True,
                                           chunk=len(ks))
    dt = time.perf_counter() - t
    return len(ks), dt, int(info["nsteps"].sum())


def run_reference(args, rank):
    if rank != 0:
        return
    tab = load_tables()
    stride = 128
    times, modes = [], 0
    for i in range(args.warmup + args.steps):
        m, dt, st = cpu_reference_pass(tab, stride)
        if i >= args.warmup:
            times.append(dt)
            modes = m
    ms = 1e3 * float(np.mean(times))
    value = modes / (ms * 1e-3)
    sample = (f"every {stride}th mode of the 512-mode workload ({modes} modes, log-spaced over the full k range) per step; "
              "restated-reference CPU (NumPy oracle: dense Jacobian + LAPACK getrf/getrs, lock-step over modes), not JAX: "
              "jax/diffrax are not installable here")
    line = dict(impl="reference", metric=METRIC, value=value, unit="k-modes/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=WORKLOAD["name"], sample=sample),
                cpu_baseline=dict(value=value, unit="k-modes/s", cores=os.cpu_count(), kind="port", sample=sample),
                e2e=dict(value=value, unit="k-modes/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][2]) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = set()
            for r in rows:
                for j, nme in enumerate(names):
                    if r[5 + j].strip().lower() == "active":
                        reasons.add(nme)
            out["reasons"] = sorted(reasons)
            out["samples"] = len(rows)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from discoeb_b200 import _cabi
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.default_library()
    tab = load_tables()
    lg, lp, lr, ln, nq = WORKLOAD["dims"]
    nk, nout = WORKLOAD["nk"], len(WORKLOAD["aexp_out"])
    n = lib.nvar(lg, lp, lr, ln, nq)
    ks = np.geomspace(WORKLOAD["kmin"], WORKLOAD["kmax"], nk)
    # weak scaling: every rank owns one cosmology's 512 modes.  Ranks > 0 perturb n_s/A_s only
    # (post-processing scalars), so all ranks do the same amount of work on distinct inputs.
    scal = tab.scalars.copy()
    scal[16] *= 1.0 + 0.01 * rank
    dims = _cabi.make_dims(ncosmo=1, nk=nk, nout=nout, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth,
                           nnu=tab.nnu, max_steps=WORKLOAD["max_steps"], power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=WORKLOAD["rtol"], atol=WORKLOAD["rtol"])
    f64 = dict(dtype=torch.float64, device=dev)
    d_sc = torch.from_numpy(scal[None]).to(dev)
    d_tb = torch.from_numpy(tab.tables[None].copy()).to(dev)
    d_k = torch.from_numpy(ks).to(dev)
    d_a = torch.tensor(WORKLOAD["aexp_out"], **f64)
    d_y = torch.zeros((1, nk, nout, 20), **f64)
    d_pk = torch.zeros((1, nk, nout), **f64)
    d_tau = torch.zeros((1, nout), **f64)
    d_st = torch.zeros((1, nk), dtype=torch.int32, device=dev)
    d_ns = torch.zeros((1, nk), dtype=torch.int32, device=dev)
    d_na = torch.zeros((1, nk), dtype=torch.int32, device=dev)
    d_ws = torch.zeros(1024, dtype=torch.int32, device=dev)       # >= deb_workspace_bytes(dims) = 256 + 8 ncosmo
    gathered = torch.zeros((world, nk, nout, 20), **f64) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)      # > 126 MB L2

    def device_step():
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.lib.deb_evolve_f64(C.byref(dims), C.byref(ctrl), d_sc.data_ptr(), d_tb.data_ptr(), d_k.data_ptr(), d_a.data_ptr(),
                                    d_y.data_ptr(), d_pk.data_ptr(), d_tau.data_ptr(), d_st.data_ptr(), d_ns.data_ptr(),
                                    d_na.data_ptr(), d_ws.data_ptr(), C.c_size_t(4096), C.c_void_p(st))
        if rc != 0:
            raise RuntimeError(lib.strerror(rc))
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_y[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    assert int(d_st.abs().max().item()) == 0, "some modes did not complete"
    total_steps = int(d_ns.sum().item())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for e0, e1 in ev:
        flush.fill_(1.0)                       # L2 flush between timed iterations (outside the event pair)
        e0.record()
        device_step()
        e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    ms_dev = sum(e0.elapsed_time(e1) for e0, e1 in ev) / args.steps
    # e2e through the host C-ABI entry (host buffers, copies inside the timed region)
    h_y = np.zeros((1, nk, nout, 20))
    e2e_ms = []
    for i in range(2 + args.steps):
        t = time.perf_counter()
        out = lib.evolve_host(dims, ctrl, scal[None], tab.tables[None], ks, np.asarray(WORKLOAD["aexp_out"]), device=local, want_pk=True)
        if i >= 2:
            e2e_ms.append(1e3 * (time.perf_counter() - t))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e = float(np.mean(e2e_ms))
    h2d = scal.nbytes + tab.tables.nbytes + ks.nbytes + 8 * nout
    d2h = h_y.nbytes + 8 * nk * nout + 8 * nout + 3 * 4 * nk
    if world > 1:
        t = torch.tensor([ms_dev, e2e], **f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, e2e = float(t[0]), float(t[1])
    if rank == 0:
        peak1, _ = lib.fp64_peak_tflops(local)
        peak = peak1 * world
        flops = world * total_steps * f_step(n)          # every rank integrates the same number of steps
        achieved = flops / (ms_dev * 1e-3) / 1e12
        line = dict(metric=METRIC, value=world * nk / (ms_dev * 1e-3), unit="k-modes/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_dev, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload=WORKLOAD["name"], modes_per_gpu=nk, attempted_steps_per_pass=total_steps,
                                l2="flushed between timed iterations (256 MB write)", parallelism=f"k-modes x{world} (independent batches, all-gather of y)"),
                    roofline=dict(bound="fp64", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                                  traffic=ncu_dram_bytes_per_launch(), kernel="k_evolve_team<3,4,2> (one CTA of 4 warps per mode)",
                                  note="FP64 FMA pipe (the path is neither HBM- nor tensor-bound); peak measured on this GPU by "
                                       "deb_fp64_peak_tflops (dependent-free DFMA streams) x n_gpus; algorithmic flops = "
                                       "(370 n + 3000) x attempted steps, n=265; traffic = dram bytes read+written per k_evolve_team "
                                       "launch from profiles/r1_v24_k_evolve_team_ncu_summary.txt (HBM is idle)"),
                    e2e=dict(value=world * nk / (e2e * 1e-3), unit="k-modes/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=e2e),
                    gpu_launches=3 * args.steps,  # k_tau_out, k_evolve_team, k_learn_order per pass (profiles/r1_v24_launch_list_bench.txt)
                    clocks=clocks)
        if world == 1 and not args.no_cpu_baseline:
            m, dt, st = cpu_reference_pass(tab, 64)
            line["cpu_baseline"] = dict(value=m / dt, unit="k-modes/s", cores=os.cpu_count(), kind="port",
                                        sample=f"every 64th mode of the workload ({m} modes, {st} attempted steps) in {dt:.1f} s; NumPy oracle "
                                               "(dense Jacobian + LAPACK LU), restated reference, not JAX")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
