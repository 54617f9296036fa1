#!/usr/bin/env python3
"""bench.py -- XC build throughput (Mgridpts/s) and per-SCF-iteration V_xc time on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one DFT_ComputeXC call (one SCF iteration's E_xc + V_xc build, what dft.py:205-208
times) on synthetic inputs of the named molecule shape: geometry from the reference's atom_txt,
STO-3G, PySCF-level-3-sized synthetic grid, seeded idempotent density matrix.
With N > 1 ranks the grid points are sharded (strong scaling: the molecule is fixed) and the
nao x nao partial V_xc and E_xc are all-reduced with NCCL inside the call.

One JSON line is printed by rank 0.  `value` is device-timed with the inputs resident in HBM;
`e2e` is the same metric through the C ABI with HOST buffers for the per-iteration inputs/outputs
(dft.py:200 uploads D, dft.py:211 downloads V_xc), copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "xc_build_mgridpts_per_s"
UNIT = "Mgridpts/s"


# --------------------------------------------------------------------------- helpers
class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs.  NVML in-process (a query costs
    microseconds); spawning nvidia-smi every 100 ms instead was seen to stall kernel launches for
    milliseconds -- with K = 2 steps that doubled a 2.4 ms step.  nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES if the launcher set it
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        act = lambda bit: "Active" if r & bit else "Not Active"
        return [str(sm), str(mx), act(n.nvmlClocksEventReasonHwSlowdown), act(n.nvmlClocksEventReasonHwThermalSlowdown),
                act(n.nvmlClocksEventReasonSwThermalSlowdown), act(n.nvmlClocksEventReasonSwPowerCap)]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    p = [x.strip() for x in out.strip().split(",")]
                    if len(p) >= 6:
                        self.samples.append(p)
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.25)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json (of measured)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback of B200_PROFILING.md (of fallback)"


def ncu_traffic(workload_key, kernel_key, world):
    """DRAM bytes per launch of a kernel from the committed ncu capture of the same command line
    (profiles/*_ncu_summary.json, written by tools/ncu_summary.py).  Captured at 1 GPU: null otherwise."""
    if world != 1:
        return None, None
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        for f in sorted(os.listdir(pdir)):
            if f.endswith("_ncu_summary.json"):
                db = json.load(open(os.path.join(pdir, f)))
                k = db.get(workload_key, {}).get(kernel_key)
                if k:
                    best = (k.get("dram_bytes"), f"profiles/{f} ({k.get('source')})")
    except Exception:
        return None, None
    return best if best else (None, None)


def cpu_port_baseline(hp, sample_points):
    """PySCF-numint-shaped CPU port (oracle/numint_port.py) on a bounded sample of the same workload."""
    from oracle import numint_port, oracle as O
    n = min(sample_points, hp.ngrid)
    idx = slice(0, n)
    xc = {"LDA": 0, "GGA": 1, "B3LYP": 2}[hp.functional]
    ao, grad = O.eval_ao(hp.coords[idx], hp.basis, deriv=1)
    w = hp.weights[idx]
    numint_port.nr_rks(xc, hp.dm, ao[: min(n, 4096)], w[: min(n, 4096)], grad[:, : min(n, 4096)])  # warm BLAS
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        numint_port.nr_rks(xc, hp.dm, ao, w, grad)
        best = min(best, time.perf_counter() - t0)
    return {"value": n / best / 1e6, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"first {n} grid points of the workload, best of 3, numpy/OpenBLAS dgemm + OpenMP pointwise "
                      f"(PySCF-numint-shaped restatement; PySCF itself is not installable here)",
            "seconds": best}


# --------------------------------------------------------------------------- shared by both arms
def needs_l2_flush(hp, n_local):
    """Inputs per rank vs L2 (126 MB): flush explicitly when they could stay (partly) cache-resident between steps.
    Inputs of more than twice the L2 are streamed through it every step (C4 on 8 GPUs: 398 MB per rank); flushing
    there only de-synchronises the ranks -- every step is then bracketed by its own device synchronise, and the
    all-reduce waits for the rank that flushed last."""
    P = 1 if hp.functional == "LDA" else 4
    return 8.0 * n_local * P * hp.nao < 2 * 126e6


def bench_config(args, hp, need_flush):
    """The `config` object -- built by ONE function so that both arms print the identical object (the driver compares
    them: `same_config`).  Everything arm-specific (engine path, rank count) lives outside it."""
    return {"workload": f"{args.workload}: {hp.name}", "functional": hp.functional, "ngrid": hp.ngrid, "nao": hp.nao,
            "basis": "sto-3g", "grid": "synthetic, PySCF level-3 point counts" if args.scale == 1.0 else
            f"synthetic, PySCF level-3 point counts x {args.scale}",
            "density": "seeded idempotent D = 2CC^T",
            "parallelism": "grid points sharded over n_gpus ranks (the reference's own code is single-GPU)",
            "l2": "explicit 384 MB flush between timed steps" if need_flush else "AO planes far larger than L2"}


def load_reference_lib():
    """ctypes handle of the reference's own CUDA (dft_solver.cu compiled UNMODIFIED for sm_100a by oracle/Makefile
    into oracle/_ref/dft_ref.so), bound exactly as dft.py:27-50 binds it; None if the prebuilt .so did not travel."""
    import ctypes
    ref_so = os.path.join(ROOT, "oracle", "_ref", "dft_ref.so")
    if not os.path.exists(ref_so):
        return None
    lib = ctypes.CDLL(ref_so)
    lib.DFT_CreateSolver.argtypes = [ctypes.c_int]; lib.DFT_CreateSolver.restype = ctypes.c_void_p
    lib.DFT_DestroySolver.argtypes = [ctypes.c_void_p]; lib.DFT_DestroySolver.restype = None
    lib.DFT_ComputeXC.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_uint64] * 5
    lib.DFT_ComputeXC.restype = ctypes.c_double
    return lib


def reference_xc(lib, functional, ngrid, nao, d_dm, d_ao, d_grad, d_w, d_vxc):
    """One call of the reference's DFT_ComputeXC on device arrays; returns (E_xc, raw V_xc on the host)."""
    from quantum_compute_dft_b200 import cuda_rt, workload
    s = lib.DFT_CreateSolver(workload.FUNCTIONAL_TYPE[functional])
    e = lib.DFT_ComputeXC(s, ngrid, nao, d_dm.data.ptr, d_ao.data.ptr, d_grad.data.ptr if d_grad is not None else 0,
                          d_w.data.ptr, d_vxc.data.ptr)
    cuda_rt.synchronize()
    v = d_vxc.get()
    lib.DFT_DestroySolver(s)
    return e, v


def parity_record(e, v, e_ref, v_ref, against):
    """E_xc / sym(V_xc) differences in BASELINE.json's terms (|dE| <= 1e-8 Ha, max |d 1/2 (V + V^T)| <= 1e-9)."""
    de = abs(float(e) - float(e_ref))
    dv = float(np.max(np.abs(0.5 * (v + v.T) - 0.5 * (v_ref + v_ref.T))))
    return {"e_xc_abs_err": de, "vxc_max_abs_err": dv, "e_xc": float(e), "e_xc_ref": float(e_ref),
            "tolerance": {"e_xc": 1e-8, "vxc": 1e-9}, "ok": bool(de <= 1e-8 and dv <= 1e-9), "against": against}


# --------------------------------------------------------------------------- reference arm
def run_reference_arm(args, rank, world):
    """The reference's own implementation of the path.  Its implementation IS CUDA
    (/root/reference/src/dft_solver.cu), compiled unmodified for sm_100a into oracle/_ref/dft_ref.so by
    oracle/Makefile; it runs on the same B200 through its own C ABI on the WHOLE grid of the workload (the same
    `config` as our arm).  Nothing of the engine is in this process: the AO planes come from the CPU oracle
    (oracle.eval_ao, chunk by chunk, uploaded with cudaMemcpy) and the device arrays from the ctypes CUDA-runtime
    shim, so the only native code on the timed path is dft_ref.so (+ cuBLAS).  If the prebuilt .so did not travel,
    the oracle's CPU port is timed on the host cores instead."""
    if rank != 0:
        return
    from oracle import oracle as O
    from quantum_compute_dft_b200 import cuda_rt, workload
    from quantum_compute_dft_b200.cuda_rt import DeviceArray
    hp = workload.host_problem(args.workload, scale=args.scale)
    n = hp.ngrid if args.ref_sample <= 0 else min(hp.ngrid, args.ref_sample)
    need_flush = needs_l2_flush(hp, n)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(args, hp, need_flush)}
    lib = load_reference_lib() if cuda_rt.device_count() > 0 else None
    if lib is not None:
        cuda_rt.set_device(0)
        nao, lda = hp.nao, hp.functional == "LDA"
        d_ao = DeviceArray((n, nao))
        d_grad = None if lda else DeviceArray((3, n, nao))
        CH = 32768
        for o in range(0, n, CH):   # AO planes from the CPU oracle, uploaded chunk by chunk
            m = min(CH, n - o)
            if lda:
                ao = O.eval_ao(hp.coords[o:o + m], hp.basis, deriv=0)
            else:
                ao, g = O.eval_ao(hp.coords[o:o + m], hp.basis, deriv=1)
                for c in range(3):
                    cuda_rt.check(cuda_rt.rt().cudaMemcpy(d_grad.data.ptr + 8 * ((c * n + o) * nao), g[c].ctypes.data,
                                                          8 * m * nao, cuda_rt.H2D), "H2D grad")
            cuda_rt.check(cuda_rt.rt().cudaMemcpy(d_ao.data.ptr + 8 * o * nao, ao.ctypes.data, 8 * m * nao, cuda_rt.H2D), "H2D ao")
        d_dm, d_w = DeviceArray.from_host(hp.dm), DeviceArray.from_host(hp.weights[:n])
        d_vxc = DeviceArray((nao, nao), zero=True)
        s = lib.DFT_CreateSolver(workload.FUNCTIONAL_TYPE[hp.functional])
        call = lambda: lib.DFT_ComputeXC(s, n, nao, d_dm.data.ptr, d_ao.data.ptr, d_grad.data.ptr if d_grad else 0,
                                         d_w.data.ptr, d_vxc.data.ptr)
        flush_buf = DeviceArray((48 * 1024 * 1024,), np.float64) if need_flush else None
        e_xc = None
        for _ in range(max(1, args.warmup)):
            e_xc = call()
        cuda_rt.synchronize()
        e0, e1 = cuda_rt.Event(), cuda_rt.Event()
        total_ms = 0.0
        if not need_flush:
            e0.record(0)
            for _ in range(args.steps):
                call()
            e1.record(0)
            e1.synchronize()
            total_ms = e0.elapsed_ms(e1)
        else:
            for _ in range(args.steps):
                flush_buf.fill_zero()
                cuda_rt.synchronize()
                e0.record(0)
                call()
                e1.record(0)
                e1.synchronize()
                total_ms += e0.elapsed_ms(e1)
        ms = total_ms / args.steps
        val = n / (ms * 1e-3) / 1e6
        sample = (f"the whole grid ({n} points) per step" if n == hp.ngrid else
                  f"first {n} of {hp.ngrid} grid points per step")
        line.update({"value": val, "ms_per_step": ms, "e_xc": e_xc,
                     "cpu_baseline": {"value": val, "unit": UNIT, "cores": 0, "kind": "reference",
                                      "sample": "reference CUDA (dft_solver.cu, unmodified, sm_100a) on the same B200, "
                                                + sample + "; inputs from the CPU oracle's AO evaluation"},
                     "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:
        cb = cpu_port_baseline(hp, args.cpu_sample)
        cb["kind"] = "port"
        line.update({"value": cb["value"], "ms_per_step": cb["seconds"] * 1e3, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", help="C1..C5 (BASELINE.json configs); default C5 = B3LYP on "
                    "C33H56N7O17P3S, the configuration the metric's target is quoted on")
    ap.add_argument("--scale", type=float, default=1.0, help="grid size factor (1.0 = PySCF level-3 point counts)")
    ap.add_argument("--cpu-sample", type=int, default=60000)
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: grid points per step (0 = the whole grid)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check of the timed inputs (sweeps, ncu runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 generic, 2 TMA")
    ap.add_argument("--devices", type=int, default=1,
                    help="single-process multi-GPU (csrc/fanout.cu): this ONE process holds the whole grid on device 0 "
                         "exactly like dft.py and DFT_ComputeXC fans out to this many GPUs; not combinable with torchrun")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="engine tuning option passed to DFT_SetOption (e.g. vxc_vk=16, vxc_shape=160, l2_prefetch=0)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    from quantum_compute_dft_b200 import cuda_rt, workload
    if cuda_rt.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    cuda_rt.set_device(local_rank)

    dist = None
    if world > 1:
        import torch.distributed as dist   # plumbing only: rendezvous, barrier, max-over-ranks
        dist.init_process_group("gloo", init_method="env://")

    hp = workload.host_problem(args.workload, scale=args.scale)
    solver = workload.make_solver(hp.functional)
    solver.set_option("path", args.path)
    solver.set_option("timing", 1)      # (per-kernel CUDA events: on for the AO evaluation and the stat-collection steps only)
    if args.devices > 1:
        if world > 1:
            raise SystemExit("--devices is the single-process mode: launch it without torchrun")
        solver.set_option("devices", args.devices)
    for kv in args.opt:
        k, v = kv.split("=")
        solver.set_option(k, float(v))
    if world > 1:
        ids = [solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        solver.comm_init(rank, world, ids[0])
    dp = workload.device_problem(hp, solver, rank, world)
    nao, n_local = dp.nao, dp.ngrid

    # subsystem (a): time the AO evaluation that produced the inputs (re-run twice, keep the faster)
    ao_ms = None
    if n_local > 0:
        for _ in range(2):
            solver.eval_ao(dp.d_coords, hp.basis, dp.d_ao, dp.d_ao_grad)
            t_ao = solver.stat("ao_ms")
            ao_ms = t_ao if ao_ms is None else min(ao_ms, t_ao)

    first_call_ms = 0.0

    def step():
        return solver.compute_xc(n_local, nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)

    def barrier():
        cuda_rt.synchronize()
        if dist is not None:
            dist.barrier()

    P = 1 if hp.functional == "LDA" else 4
    need_flush = needs_l2_flush(hp, n_local)
    flush_buf = cuda_rt.DeviceArray((48 * 1024 * 1024,), np.float64) if need_flush else None  # 384 MB

    # W untimed warm-up steps; the last two of them run AFTER the barrier that opens the timed region (below), so that the
    # K timed steps start on a stream that is already busy: a sub-millisecond step (C4 on 8 GPUs: 0.5 ms) timed right
    # behind a host-side rendezvous measured 0.85 ms against 0.55 ms in the end-to-end loop that follows it
    late_warm = 2 if args.warmup >= 3 else 0
    t_first = time.perf_counter()
    for i in range(args.warmup - late_warm):
        step()
        if i == 0:
            first_call_ms = (time.perf_counter() - t_first) * 1e3    # fan-out: includes cutting the resident shards
    e_xc = step() if late_warm == 0 else None

    # ---- device-timed region: K steps, CUDA events on the engine's stream
    stream = solver.stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = cuda_rt.Event(), cuda_rt.Event()
    dens_ms = vxc_ms = 0.0
    solver.set_option("timing", 0)      # the timed regions run the library as a caller would: no per-kernel events
    barrier()
    for _ in range(late_warm):
        e_xc = step()
    if not need_flush:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        ev1.synchronize()
        total_ms = ev0.elapsed_ms(ev1)
    else:
        total_ms = 0.0
        for _ in range(args.steps):
            flush_buf.fill_zero()          # evict L2 (384 MB > 126 MB), outside the timed events
            cuda_rt.synchronize()
            ev0.record(stream)
            step()
            ev1.record(stream)
            ev1.synchronize()
            total_ms += ev0.elapsed_ms(ev1)
    # per-kernel durations (the engine's own CUDA events on its stream) for the roofline: read in K further steps on the
    # same inputs right after the timed region -- recording and reading them costs a handful of driver calls and event
    # records per step, which inside the region would be charged to the step (a tenth of it at H2O size; visible at
    # 8 GPUs, where a C4 step is 0.5 ms)
    solver.set_option("timing", 1)
    for _ in range(args.steps):
        if need_flush:
            flush_buf.fill_zero()
            cuda_rt.synchronize()
        step()
        dens_ms += solver.stat("density_ms"); vxc_ms += solver.stat("vxc_ms")
    solver.set_option("timing", 0)
    barrier()

    # ---- end-to-end region: host buffers for the per-iteration input (D) and outputs (V_xc, E_xc)
    h_dm = cuda_rt.PinnedArray((nao, nao)); h_dm.array[...] = hp.dm
    h_v = cuda_rt.PinnedArray((nao, nao))
    nbytes = nao * nao * 8
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cuda_rt.memcpy_async(dp.d_dm.data.ptr, h_dm.ptr, nbytes, cuda_rt.H2D, stream)   # dft.py:200
        e_host = step()                                                                  # dft.py:206 (returns E_xc)
        cuda_rt.memcpy_async(h_v.ptr, dp.d_vxc.data.ptr, nbytes, cuda_rt.D2H, stream)   # dft.py:211
        cuda_rt.stream_synchronize(stream)
    cuda_rt.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    times = np.array([total_ms, e2e_ms, dens_ms, vxc_ms], dtype=np.float64)
    if dist is not None:
        import torch
        t = torch.from_numpy(times)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times = t.numpy()
    total_ms, e2e_ms, dens_ms, vxc_ms = (float(x) for x in times)

    # ---- parity of exactly what was timed (outside every timed region): the result of the last step against the
    # reference's own CUDA on the same inputs and, at N > 1, against a 1-GPU recomputation by this engine on rank 0
    parity = None
    if not args.no_parity:
        v_timed = dp.d_vxc.get()           # after the all-reduce: every rank holds the global matrix
        ranks_diff = 0.0
        if dist is not None:               # ... and they must all hold the SAME one
            import torch
            t0_ = torch.from_numpy(v_timed.copy())
            dist.broadcast(t0_, src=0)
            td = torch.tensor([float(np.max(np.abs(v_timed - t0_.numpy())))], dtype=torch.float64)
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            ranks_diff = float(td[0])
        if rank == 0:
            if world > 1:
                solver1 = workload.make_solver(hp.functional)
                solver1.set_option("path", args.path)
                dp_full = workload.device_problem(hp, solver1)
                e1 = solver1.compute_xc(dp_full.ngrid, nao, dp_full.d_dm, dp_full.d_ao, dp_full.d_weights, dp_full.d_vxc,
                                        dp_full.d_ao_grad)
                v1 = dp_full.d_vxc.get()
                single = parity_record(e_host, v_timed, e1, v1, "this engine on 1 GPU (rank 0), whole grid")
            else:
                dp_full, single = dp, None
            ref = load_reference_lib()
            if ref is not None:
                d_vref = cuda_rt.DeviceArray((nao, nao), zero=True)
                e_r, v_r = reference_xc(ref, hp.functional, dp_full.ngrid, nao, dp_full.d_dm, dp_full.d_ao, dp_full.d_ao_grad,
                                        dp_full.d_weights, d_vref)
                parity = parity_record(e_host, v_timed, e_r, v_r,
                                       "reference CUDA (dft_solver.cu unmodified, oracle/_ref/dft_ref.so) on the same "
                                       "device arrays, whole grid, one GPU")
            else:
                parity = {"against": None, "ok": None, "note": "oracle/_ref/dft_ref.so did not travel"}
            if single is not None:
                parity["vs_single_gpu"] = single
                parity["ranks_max_abs_diff"] = ranks_diff
                parity["ok"] = bool(parity.get("ok") in (True, None) and single["ok"] and ranks_diff <= 1e-12)
        barrier()

    if rank == 0:
        K = args.steps
        ms_step = total_ms / K
        value = hp.ngrid / (ms_step * 1e-3) / 1e6
        e2e_value = hp.ngrid / (e2e_ms / K * 1e-3) / 1e6
        peaks, peak_src = measured_peaks()
        fanned = int(solver.stat("fan_active")) == 1
        ndev = args.devices if fanned else 1
        flops_local = workload.algorithmic_flops(n_local, nao) / ndev    # (kernel times are the slowest device's)
        dens_avg, vxc_avg = dens_ms / K, vxc_ms / K
        tensor_bound = nao >= 100      # crossover of FP64-tensor and HBM rooflines near nao ~ 100
        if tensor_bound:
            dom, dom_ms = ("vxc_kernel", vxc_avg) if vxc_avg >= dens_avg else ("density_kernel", dens_avg)
            dmma_peak = solver.lib.DFT_MicrobenchDMMA(4096)
            dfma_peak = solver.lib.DFT_MicrobenchDFMA(4096)
            achieved = 0.5 * flops_local / (dom_ms * 1e-3) / 1e12
            traffic, traffic_src = ncu_traffic(args.workload, dom.split("_")[0], world)
            roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                        "frac": achieved / dmma_peak if dmma_peak > 0 else None, "traffic": traffic,
                        "traffic_source": traffic_src,
                        "peak_source": "DFT_MicrobenchDMMA measured in this run (register-resident mma.sync m8n8k4 f64 "
                                       "-> DMMA); MEASURED_PEAKS.json has no FP64 entry; nominal 37-40 TFLOP/s",
                        "dfma_peak_tflops": dfma_peak,
                        "whole_call_tflops": flops_local / (ms_step * 1e-3) / 1e12,
                        "density_ms": dens_avg, "vxc_ms": vxc_avg,
                        "algorithmic_flops_per_launch": 0.5 * flops_local}
            # the same with the DMMAs the kernel actually executed (AO screening skips exact-zero fragments): comparable
            # with ncu's FP64-tensor-pipe utilisation of that kernel
            skipped = solver.stat("vxc_skip_fraction") if dom == "vxc_kernel" else solver.stat("skip_fraction")
            if roofline["frac"] is not None and skipped is not None and 0.0 <= skipped < 1.0:
                roofline["executed_fraction_of_dense_work"] = 1.0 - skipped
                roofline["frac_executed"] = roofline["frac"] * (1.0 - skipped)
        else:
            bytes_local = workload.algorithmic_bytes(n_local, nao, hp.functional)
            achieved = bytes_local / (ms_step * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "whole call (density + vxc kernels)", "achieved": achieved,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                        "peak_source": peak_src, "density_ms": dens_avg, "vxc_ms": vxc_avg,
                        "algorithmic_bytes_per_call": bytes_local}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_port_baseline(hp, args.cpu_sample)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world * ndev, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, hp, need_flush),
            "engine": {"path": int(solver.stat("path")), "ranks": world, "points_per_rank": n_local // ndev,
                       "input_gb_per_rank": 8.0 * n_local * P * nao / 1e9 / ndev, "options": args.opt,
                       "process_model": "one process per GPU, NCCL all-reduce" if world > 1 else
                       ("ONE process, DFT_ComputeXC fans out to %d devices (all arrays on device 0 as in dft.py); shards cut "
                        "%d time(s), first call %.1f ms" % (ndev, int(solver.stat("fan_scatters")), first_call_ms)
                        if fanned else "one process, one GPU")},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 8,
                    "note": "per SCF iteration the driver uploads D (dft.py:200) and downloads V_xc (dft.py:211) + E_xc; "
                            "AO planes stay resident across iterations as in dft.py:155-176"},
            "gpu_launches": int(solver.stat("launches")) * K,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary(),
            "per_scf_iter_vxc_ms": ms_step, "e_xc": e_host, "parity": parity,
            "ao_screening": {"density_ksteps_skipped_frac": solver.stat("skip_fraction"),
                             "vxc_fragments_skipped_frac": solver.stat("vxc_skip_fraction"),
                             "note": "share of the density kernel's 32x4 Phi fragments / of the V kernel's (8-column B "
                                     "fragment, k-step) units that are exact zeros and skipped (rank 0); "
                                     "roofline.achieved stays quoted on the DENSE flop count"},
            "ao_eval": None if not ao_ms else {
                "ms": ao_ms, "bytes_written": 8.0 * n_local * nao * P,
                "achieved_gbs": 8.0 * n_local * nao * P / (ao_ms * 1e-3) / 1e9,
                "frac_of_hbm": 8.0 * n_local * nao * P / (ao_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "note": "DFT_EvalAO (subsystem a), once per SCF run, rank 0's slice; not part of the timed step"},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        solver.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
