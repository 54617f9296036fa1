#!/usr/bin/env python3
"""Time DFT_EvalAO for one workload with each block shape.  Usage: python tools/ao_variants.py C5"""
import sys
sys.path.insert(0, ".")
from quantum_compute_dft_b200 import workload
from quantum_compute_dft_b200.solver import DFTSolverWrapper, DEFAULT_LIB
hp = workload.host_problem(sys.argv[1])
s = DFTSolverWrapper(DEFAULT_LIB, hp.functional)
dp = workload.device_problem(hp, s)
for shape in (0, 16, 32):
    s.set_option("ao_shape", shape)
    best = 1e9
    for _ in range(3):
        s.eval_ao(dp.d_coords, hp.basis, dp.d_ao, dp.d_ao_grad); best = min(best, s.stat("ao_ms"))
    P = 4 if dp.d_ao_grad is not None else 1
    print(sys.argv[1], "ao_shape", shape, "ao_ms %.4f  %.0f GB/s" % (best, 8.0 * dp.ngrid * dp.nao * P / best / 1e6), flush=True)
