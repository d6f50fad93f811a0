#!/usr/bin/env python
"""bench.py -- headline benchmark of the Einstein-Boltzmann hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic input: every k-mode of
BASELINE.json configs[1] (LambdaCDM + one massive neutrino, lmax=31 -> 32 multipoles per
hierarchy, nq=5, n=265 equations per mode, num_k=512, k in [1e-4, 10]/Mpc, a_out=1,
rtol=atol=1e-4, reference solver defaults) integrated from its start time to z=0 and converted
to the 20 output fields + P(k).  Inputs are the committed fiducial tables (tests/golden).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU): the step integrates ONE k grid of 512 x N modes over the same k range,
dealt round-robin over the ranks (512 modes per GPU: weak scaling, no data-path collective during the solve), and the
library's own multi-GPU entry (deb_evolve_sharded_f64: ncclAllGather on the device) leaves the full result on every
rank inside the timed region.  Sub-records (key `sub`) time BASELINE config 3 (4096 modes dealt over the N ranks: strong
scaling) and config 4 (128 default_rng(0) cosmologies x 256 k per GPU, tables produced on the GPU first).

`value`   : k-modes/s, whole job, inputs resident in HBM, CUDA-event time of the step (max over ranks).
`e2e`     : same metric through the host C-ABI entry (deb_evolve_host_f64) with HOST buffers:
            H2D of scalars/tables/k, kernel, D2H of y/P(k)/status inside the timed region.
`roofline`: FP64-FMA pipe (this path is neither HBM- nor tensor-bound): algorithmic flops
            F_step(n) = 370 n + 3000 per attempted step x attempted steps,
            over the kernel's CUDA-event time, against the DFMA peak measured on this GPU by
            deb_fp64_peak_tflops (MEASURED_PEAKS.json holds no FP64 figure).
`cpu_baseline`: the restated reference (NumPy/LAPACK oracle, dense Jacobian + dense LU, lock-step
            over modes like the reference's vmap) on the host cores, on a bounded sub-sample; `cpu_structured` next to
            it = the kernel SOURCE compiled for the host (tests/emu, OpenMP, all cores) on all 512 modes: what the same
            structured algorithm does on the CPU, i.e. what the B200 buys.
`parity`  : the timed step's own P(k) against the oracle's committed vector for all 512 modes
            (tests/golden/oracle_config2_full512.npz): fraction within 1e-5 and the worst deviation.
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


# dram__bytes_read.sum + dram__bytes_write.sum of one k_evolve_duo<3> launch on this workload (ncu --set full),
# read from the committed summary so that the number and its evidence cannot drift apart
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_final_k_duo_ncu_summary.txt")


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
    """One bounded pass: every `stride`-th mode of the workload through the NumPy oracle, ALL of them in one
    lock-step chunk (the way jax.vmap advances the reference's modes, perturbations.py:980-987)."""
    import oracle.discoeb_oracle as O
    p = tab.param()
    ks = np.geomspace(WORKLOAD["kmin"], WORKLOAD["kmax"], WORKLOAD["nk"])[stride - 1::stride]
    lg, lp, lr, ln, nq = WORKLOAD["dims"]
    t = time.perf_counter()
    y, k, _, info = O.evolve_perturbations(param=p, aexp_out=WORKLOAD["aexp_out"], kmin=0, kmax=0, num_k=len(ks), kmodes=ks,
                                           lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, rtol=WORKLOAD["rtol"],
                                           atol=WORKLOAD["rtol"], max_steps=WORKLOAD["max_steps"], return_info=True,
                                           chunk=len(ks))
    dt = time.perf_counter() - t
    return len(ks), dt, int(info["nsteps"].sum())


def cpu_structured_pass(tab):
    """The kernel source compiled for the host (tests/emu/_build/libdeb_emu.so, OpenMP over modes) on the whole workload."""
    from discoeb_b200 import _cabi
    path = os.path.join(ROOT, "tests", "emu", "_build", "libdeb_emu.so")
    if not os.path.exists(path):
        return None
    lib = _cabi.Library(path, prefix="emu_")
    lg, lp, lr, ln, nq = WORKLOAD["dims"]
    ks = np.geomspace(WORKLOAD["kmin"], WORKLOAD["kmax"], WORKLOAD["nk"])
    dims = _cabi.make_dims(ncosmo=1, nk=len(ks), nout=1, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth, nnu=tab.nnu,
                           max_steps=WORKLOAD["max_steps"], power_idx=4)
    ctrl = _cabi.make_ctrl(rtol=WORKLOAD["rtol"], atol=WORKLOAD["rtol"])
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        lib.evolve_host(dims, ctrl, tab.scalars[None], tab.tables[None], ks, np.asarray(WORKLOAD["aexp_out"]), want_pk=True)
        best = min(best, time.perf_counter() - t)
    return dict(value=len(ks) / best, unit="k-modes/s", cores=os.cpu_count(), kind="kernel source compiled for the host (tests/emu), OpenMP over modes",
                sample=f"all {len(ks)} modes of the workload, best of 3 ({best * 1e3:.0f} ms per pass)")


def run_reference(args, rank):
    """Reference arm: the restated reference on the host cores.  A step = one lock-step pass over a bounded sample of
    the workload; the sample is the largest of 64 / 32 / 16 modes that keeps warm-up + K steps within ~5 minutes
    (calibrated on a first 16-mode pass; the oracle is LAPACK-bound, so modes/s barely moves with the sample)."""
    if rank != 0:
        return
    tab = load_tables()
    m, dt, _ = cpu_reference_pass(tab, 32)
    budget = 300.0 / max(1, args.warmup + args.steps)
    stride = 8 if 4.5 * dt < budget else (16 if 2.2 * dt < budget else 32)
    times, modes = [], 0
    for i in range(args.warmup + args.steps):
        m, dt, st = cpu_reference_pass(tab, stride)
        if i >= args.warmup:
            times.append(dt)
            modes = m
    ms = 1e3 * float(np.mean(times))
    value = modes / (ms * 1e-3)
    sample = (f"every {stride}th mode of the 512-mode workload ({modes} modes, log-spaced over the full k range, ONE lock-step chunk) per step; "
              "restated-reference CPU (NumPy oracle: dense Jacobian + LAPACK getrf/getrs, lock-step over modes), not JAX: "
              "jax/diffrax are not installable here")
    line = dict(impl="reference", metric=METRIC, value=value, unit="k-modes/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=WORKLOAD["name"], sample=sample),
                cpu_baseline=dict(value=value, unit="k-modes/s", cores=os.cpu_count(), kind="port", sample=sample),
                e2e=dict(value=value, unit="k-modes/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    cs = cpu_structured_pass(tab)
    if cs:
        line["cpu_structured"] = cs
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
    from discoeb_b200.background import config4_draws, pack_background_input
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _cabi.default_library()
    L = lib.lib
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from discoeb_b200.distributed import NativeComm

        def bcast(payload):
            box = [payload]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = NativeComm(world, rank, bcast, device=local, lib=lib)      # the library's own NCCL communicator
    L.deb_sharded_workspace_bytes.restype = C.c_size_t
    L.deb_sharded_workspace_bytes.argtypes = [C.POINTER(_cabi.DebDims), C.c_int32]
    L.deb_evolve_sharded_f64.restype = C.c_int
    L.deb_evolve_sharded_f64.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(_cabi.DebDims), C.POINTER(_cabi.DebCtrl)] + [C.c_void_p] * 9 + \
        [C.c_void_p, C.c_size_t, C.c_int32] + [C.c_void_p] * 4 + [C.c_void_p]
    L.deb_background_f64.restype = C.c_int
    L.deb_background_f64.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.deb_background_workspace_bytes.restype = C.c_size_t
    L.deb_background_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    tab = load_tables()
    lg, lp, lr, ln, nq = WORKLOAD["dims"]
    nk1, nout = WORKLOAD["nk"], len(WORKLOAD["aexp_out"])
    n = lib.nvar(lg, lp, lr, ln, nq)
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    ctrl = _cabi.make_ctrl(rtol=WORKLOAD["rtol"], atol=WORKLOAD["rtol"])
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Sharded:
        """One k grid of nk modes (ncosmo cosmologies) dealt over the ranks through deb_evolve_sharded_f64 (gather = 0)."""

        def __init__(self, scal, tabs, nk, max_steps):
            self.nc, self.nk = scal.shape[0], nk
            self.ks = np.geomspace(WORKLOAD["kmin"], WORKLOAD["kmax"], nk)
            self.dims = _cabi.make_dims(ncosmo=self.nc, nk=nk, nout=nout, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=tab.nth,
                                        nnu=tab.nnu, max_steps=max_steps, power_idx=4)
            self.d_sc = scal if torch.is_tensor(scal) else torch.from_numpy(np.ascontiguousarray(scal)).to(dev)
            self.d_tb = tabs if torch.is_tensor(tabs) else torch.from_numpy(np.ascontiguousarray(tabs)).to(dev)
            self.d_k = torch.from_numpy(self.ks).to(dev)
            self.d_a = torch.tensor(WORKLOAD["aexp_out"], **f64)
            self.d_y = torch.zeros((self.nc, nk, nout, 20), **f64); self.d_pk = torch.zeros((self.nc, nk, nout), **f64)
            self.d_tau = torch.zeros((self.nc, nout), **f64)
            self.d_st = torch.zeros((self.nc, nk), **i32); self.d_ns = torch.zeros((self.nc, nk), **i32)
            self.wsb = L.deb_sharded_workspace_bytes(C.byref(self.dims), world)
            self.d_ws = torch.zeros(self.wsb // 8 + 8, **f64)

        def step(self):
            rc = L.deb_evolve_sharded_f64(comm.handle if comm else None, world, rank, C.byref(self.dims), C.byref(ctrl), self.d_sc.data_ptr(),
                                          self.d_tb.data_ptr(), self.d_k.data_ptr(), self.d_a.data_ptr(), self.d_y.data_ptr(), self.d_pk.data_ptr(),
                                          self.d_tau.data_ptr(), self.d_st.data_ptr(), self.d_ns.data_ptr(), self.d_ws.data_ptr(), C.c_size_t(self.wsb),
                                          0, None, None, None, None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc != 0:
                raise RuntimeError(lib.strerror(rc))

        def timed(self, steps, warmup):
            for _ in range(warmup):
                self.step()
            barrier()
            assert int(self.d_st.abs().max().item()) == 0, "some modes did not complete"
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            for e0, e1 in ev:
                flush.fill_(1.0)                   # L2 flush between timed iterations (outside the event pair)
                e0.record(); self.step(); e1.record()
            barrier()
            ms = sum(e0.elapsed_time(e1) for e0, e1 in ev) / steps
            if world > 1:
                t = torch.tensor([ms], **f64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            return ms

    # ---------------- headline: 512 modes per GPU (one k grid of 512 x N modes dealt over the ranks) ----------------
    head = Sharded(tab.scalars[None], tab.tables[None], nk1 * world, WORKLOAD["max_steps"])
    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        head.step()
    barrier()
    if rank == 0:
        sampler.start()
    ms_dev = head.timed(args.steps, 0)
    total_steps = int(head.d_ns.sum().item())
    # parity of the timed step's own output (N = 1: the 512-mode grid the committed oracle vector is on)
    parity = None
    if world == 1:
        try:
            import helpers
            ora = helpers.load_case("config2_full512")
            rel = np.abs(head.d_pk[0, :, 0].cpu().numpy() / ora["pk4"][:, 0] - 1)
            parity = dict(against="NumPy oracle, all 512 modes, tests/golden/oracle_config2_full512.npz (P(k) of delta_m at z=0)",
                          frac_within_1e5=float((rel < 1e-5).mean()), max_rel=float(rel.max()), median_rel=float(np.median(rel)),
                          same_step_counts=float((head.d_ns[0].cpu().numpy() == ora["nsteps"]).mean()),
                          note="free-running adaptive solves at rtol=1e-4: modes whose accept/reject sequence differs from the oracle's "
                               "differ by the solver's own tolerance-level error (tests/parity_checks.py: check_convergence); replay "
                               "along the reference's step sequence agrees to <= 2e-8 (tests/test_gpu_parity.py)")
        except Exception as e:      # noqa: BLE001
            parity = dict(error=repr(e)[:200])
    # e2e through the host C-ABI entry (host buffers, copies -- and at N > 1 the NCCL gather -- inside the timed region)
    ks_all = head.ks
    e2e_ms = []
    for i in range(2 + args.steps):
        barrier()
        t = time.perf_counter()
        if world == 1:
            out = lib.evolve_host(head.dims, ctrl, tab.scalars[None], tab.tables[None], ks_all, np.asarray(WORKLOAD["aexp_out"]), device=local, want_pk=True)
        else:
            out = lib.evolve_sharded_host(comm, head.dims, ctrl, tab.scalars[None], tab.tables[None], ks_all, np.asarray(WORKLOAD["aexp_out"]), want_pk=True)
        if i >= 2:
            e2e_ms.append(1e3 * (time.perf_counter() - t))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e = float(np.mean(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e], **f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t[0])
    nk_tot = nk1 * world
    h2d = tab.scalars.nbytes + tab.tables.nbytes + ks_all.nbytes + 8 * nout
    d2h = nk_tot * nout * 20 * 8 + 8 * nk_tot * nout + 8 * nout + 2 * 4 * nk_tot

    # ---------------- sub-records: BASELINE config 3 (strong scaling) and config 4 (cosmology batch) ----------------
    sub = {}
    if not args.no_sub:
        import helpers
        w0 = helpers.load_tables("w0wa")
        c3 = Sharded(w0.scalars[None], w0.tables[None], 4096, 4096)
        ms3 = c3.timed(3, 2)
        sub["config3_strong"] = dict(workload="w0wa + massive nu, n=265, 4096 k dealt round-robin over the ranks, all-gather by the library",
                                     modes=4096, ms_per_step=ms3, value=4096 / (ms3 * 1e-3), unit="k-modes/s", scaling="strong",
                                     attempted_steps=int(c3.d_ns.sum().item()))
        del c3
        # config 4: every rank owns 128 of the 1024 default_rng(0) cosmologies x 256 k; tables produced on the GPU first
        ncs = 128
        base = dict(Omegam=0.3099, Omegab=0.0488911, w_DE_0=-0.99, w_DE_a=0.0, cs2_DE=1.0, Omegak=0.0, A_s=2.1064e-09, n_s=0.96822,
                    H0=67.742, Tcmb=2.7255, YHe=0.248, Neff=2.046, Nmnu=1, mnu=0.06, k_p=0.05)
        draws = config4_draws(1024)[rank * ncs:(rank + 1) * ncs]
        d_in = torch.from_numpy(np.stack([pack_background_input({**base, **d}) for d in draws])).to(dev)
        tl = 3 * (5 * 256 + 2 * 512)
        d_sc4 = torch.zeros((ncs, 24), **f64); d_tb4 = torch.zeros((ncs, tl), **f64)
        bws = L.deb_background_workspace_bytes(ncs, 256)
        d_bws = torch.zeros(bws // 8 + 8, **f64)
        tb_ms = []
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.deb_background_f64(ncs, 256, d_in.data_ptr(), d_sc4.data_ptr(), d_tb4.data_ptr(), d_bws.data_ptr(), C.c_size_t(bws),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, lib.strerror(rc)
            tb_ms.append(e0.elapsed_time(e1))
        # solve: local cosmologies only (cosmology batches shard contiguously, no gather needed beyond concatenation)
        ks4 = np.geomspace(WORKLOAD["kmin"], WORKLOAD["kmax"], 256)
        dims4 = _cabi.make_dims(ncosmo=ncs, nk=256, nout=nout, lmaxg=lg, lmaxgp=lp, lmaxr=lr, lmaxnu=ln, nqmax=nq, nth=256, nnu=512, max_steps=4096, power_idx=4)
        d_k4 = torch.from_numpy(ks4).to(dev); d_a4 = torch.tensor(WORKLOAD["aexp_out"], **f64)
        d_y4 = torch.zeros((ncs, 256, nout, 20), **f64); d_pk4 = torch.zeros((ncs, 256, nout), **f64); d_tau4 = torch.zeros((ncs, nout), **f64)
        d_st4 = torch.zeros((ncs, 256), **i32); d_ns4 = torch.zeros((ncs, 256), **i32); d_na4 = torch.zeros((ncs, 256), **i32)
        L.deb_workspace_bytes.restype = C.c_size_t
        wsb4 = L.deb_workspace_bytes(C.byref(dims4))
        d_ws4 = torch.zeros(wsb4 // 8 + 8, **f64)
        sv_ms = []
        for i in range(3):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.deb_evolve_f64(C.byref(dims4), C.byref(ctrl), d_sc4.data_ptr(), d_tb4.data_ptr(), d_k4.data_ptr(), d_a4.data_ptr(), d_y4.data_ptr(),
                                  d_pk4.data_ptr(), d_tau4.data_ptr(), d_st4.data_ptr(), d_ns4.data_ptr(), d_na4.data_ptr(), d_ws4.data_ptr(),
                                  C.c_size_t(wsb4), C.c_void_p(torch.cuda.current_stream().cuda_stream))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, lib.strerror(rc)
            if i > 0:
                sv_ms.append(e0.elapsed_time(e1))
        ok4 = int(d_st4.abs().max().item()) == 0
        t4 = torch.tensor([min(sv_ms), min(tb_ms[1:]), float(d_ns4.sum().item()), 0.0 if ok4 else 1.0], **f64)
        if world > 1:
            tm = t4.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t4.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            t4 = torch.stack([tm[0], tm[1], ts[2], tm[3]])
        modes4 = ncs * 256 * world
        sub["config4"] = dict(workload=f"{ncs * world} numpy.random.default_rng(0) cosmologies x 256 k, n=265 ({ncs} distinct table sets per GPU, produced "
                                       "on the GPU by deb_background_f64); cosmologies sharded contiguously",
                              modes=modes4, solve_ms=float(t4[0]), tables_ms=float(t4[1]), value=modes4 / (float(t4[0]) * 1e-3),
                              value_with_tables=modes4 / ((float(t4[0]) + float(t4[1])) * 1e-3), unit="k-modes/s", scaling="weak",
                              attempted_steps=int(t4[2]), all_modes_ok=bool(float(t4[3]) == 0.0))

    if rank == 0:
        peak1, _ = lib.fp64_peak_tflops(local)
        peak = peak1 * world
        flops = total_steps * f_step(n)
        achieved = flops / (ms_dev * 1e-3) / 1e12
        launches = 3 if world == 1 else 6
        line = dict(metric=METRIC, value=nk_tot / (ms_dev * 1e-3), unit="k-modes/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_dev, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload=WORKLOAD["name"], modes_per_gpu=nk1, attempted_steps_per_pass=total_steps,
                                l2="flushed between timed iterations (256 MB write)",
                                parallelism=(f"one k grid of {nk_tot} modes dealt round-robin over {world} GPUs (512 per GPU), ncclAllGather by the library "
                                             "(deb_evolve_sharded_f64)" if world > 1 else "1 GPU"),
                                state="steady state: the caller's workspace holds the work list learned from the previous call (DESIGN.md section 3)"),
                    roofline=dict(bound="fp64", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                                  traffic=ncu_dram_bytes_per_launch(), kernel="k_evolve_duo<3> (one 4-warp team per mode, two teams per CTA, one CTA per SM)",
                                  note="FP64 FMA pipe (the path is neither HBM- nor tensor-bound); peak measured on this GPU by "
                                       "deb_fp64_peak_tflops (dependent-free DFMA streams) x n_gpus; algorithmic flops = "
                                       "(370 n + 3000) x attempted steps, n=265; traffic = dram bytes read+written per k_evolve_duo "
                                       f"launch from {os.path.relpath(NCU_SUMMARY, ROOT)} (HBM is idle)"),
                    e2e=dict(value=nk_tot / (e2e * 1e-3), unit="k-modes/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=e2e),
                    gpu_launches=launches * args.steps,
                    gpu_launches_note=("k_tau_out, k_evolve_duo, k_learn_order per pass" if world == 1 else
                                       "k_take_modes, k_tau_out, k_evolve_duo, k_learn_order, k_pack_rows, k_unpack_rows per pass per rank (+ NCCL's all-gather kernel)"),
                    clocks=clocks)
        if parity is not None:
            line["parity"] = parity
        if sub:
            line["sub"] = sub
        if world == 1 and not args.no_cpu_baseline:
            m, dt, st = cpu_reference_pass(tab, 32)
            line["cpu_baseline"] = dict(value=m / dt, unit="k-modes/s", cores=os.cpu_count(), kind="port",
                                        sample=f"every 32nd mode of the workload ({m} modes in one lock-step chunk, {st} attempted steps) in {dt:.1f} s; NumPy oracle "
                                               "(dense Jacobian + LAPACK LU), restated reference, not JAX")
            cs = cpu_structured_pass(tab)
            if cs:
                line["cpu_structured"] = cs
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the config-3 / config-4 sub-records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
