"""Time DFT_ComputeXC for several option sets on ONE device problem (built once: the C5 inputs take longer to
generate than to integrate).  Usage:  python tools/vxc_sweep.py C5 "vxc_skip=1,vxc_skip_mode=2" "vxc_skip=0" ...
Prints per option set: step / density / V milliseconds (CUDA events inside the engine, best of the timed steps
and mean), E_xc, the V kernel's skipped share, and max |V - V_first| as a cross-check between instances."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_compute_dft_b200 import workload  # noqa: E402

DEFAULTS = {"vxc_skip": -1, "vxc_skip_mode": 2, "vxc_vk": 0, "vxc_shape": 0, "zero_skip": 1, "tma_3d": 1,
            "vxc_scatter": 1, "vxc_producers": 2, "wait_ns": 0, "dyn_sched": 1, "l2_prefetch": 0, "stagger_min": 8, "density_unit": 0, "vxc_rebalance": 1, "vxc_prefetch": 0, "density_producers": 1, "density_scatter": 0, "density_wide": 0}


def main():
    wl = sys.argv[1]
    sets = sys.argv[2:] or [""]
    steps = int(os.environ.get("SWEEP_STEPS", "6"))
    hp = workload.host_problem(wl)
    # SWEEP_LIB=diag: the -DDFT_DIAGNOSTICS build (python -m quantum_compute_dft_b200.build --diag), which alone
    # accepts debug_nodmma=1 (operand-delivery floor; results are wrong)
    lib_tag = os.environ.get("SWEEP_LIB", "")
    lib_path = None
    if lib_tag:
        from quantum_compute_dft_b200.solver import DEFAULT_LIB
        lib_path = DEFAULT_LIB.replace(".so", f"_{lib_tag}.so")
    solver = workload.make_solver(hp.functional, lib_path)
    solver.set_option("timing", 1)
    # SWEEP_RANKS=N: time rank 0's shard of an N-rank run (the per-GPU problem of a multi-GPU step) on one GPU
    nranks = int(os.environ.get("SWEEP_RANKS", "1"))
    dp = workload.device_problem(hp, solver, 0, nranks)
    v_first = None
    for spec in sets:
        opts = dict(DEFAULTS)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            opts[k] = float(v)
        for k, v in opts.items():
            solver.set_option(k, v)
        rec = []
        for i in range(steps + 2):
            e = solver.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
            if i >= 2:
                rec.append((solver.stat("total_ms"), solver.stat("density_ms"), solver.stat("vxc_ms")))
        rec = np.array(rec)
        v = dp.d_vxc.get()
        if v_first is None:
            v_first = v
        print(f"{wl} shard 1/{nranks} ({dp.ngrid} pts) [{spec or 'defaults'}] step {rec[:, 0].mean():.3f} (min {rec[:, 0].min():.3f}) ms  "
              f"density {rec[:, 1].mean():.3f}  V {rec[:, 2].mean():.3f} (min {rec[:, 2].min():.3f})  "
              f"E {e!r}  v_skipped {solver.stat('vxc_skip_fraction'):.3f}  d_skipped {solver.stat('skip_fraction'):.3f}  "
              f"max|dV| {np.abs(v - v_first).max():.2e}", flush=True)
    dp.free()


if __name__ == "__main__":
    main()
