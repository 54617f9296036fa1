#!/usr/bin/env python3
"""Per-call time breakdown of DFT_ComputeXC for one workload: engine events vs wall clock.
Usage: python tools/call_breakdown.py C4 [KEY=VALUE ...]"""
import sys, time
sys.path.insert(0, ".")
from quantum_compute_dft_b200 import workload, cuda_rt
hp = workload.host_problem(sys.argv[1])
s = workload.make_solver(hp.functional)
for kv in sys.argv[2:]:
    k, v = kv.split("="); s.set_option(k, float(v))
dp = workload.device_problem(hp, s)
for i in range(8):
    cuda_rt.synchronize()
    t0 = time.perf_counter()
    e = s.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"call {i}: wall {wall:.3f} ms  density+point {s.stat('density_ms'):.3f}  vxc {s.stat('vxc_ms'):.3f}  "
          f"finalize.. {s.stat('reduce_ms'):.3f}  total {s.stat('total_ms'):.3f}  plans {int(s.stat('plans_built'))}")
