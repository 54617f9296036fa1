#!/usr/bin/env python3
"""DFT_ComputeCoulomb (J = ERI . vec(D)) bandwidth at the config molecules' nao.
The ERI is zero-filled device memory (bandwidth does not depend on the values); every ERI element is read
once: 8 nao^4 bytes.  Usage: python tools/coulomb_bench.py [nao ...]"""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from quantum_compute_dft_b200 import cuda_rt
from quantum_compute_dft_b200.cuda_rt import DeviceArray
from quantum_compute_dft_b200.solver import DFTSolverWrapper, DEFAULT_LIB

peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
s = DFTSolverWrapper(DEFAULT_LIB, "LDA")
for nao in [int(a) for a in sys.argv[1:]] or [7, 36, 152]:
    n2 = nao * nao
    eri = DeviceArray((n2, n2), zero=True)
    dm = DeviceArray.from_host(np.random.default_rng(0).standard_normal((nao, nao)))
    J = DeviceArray((nao, nao), zero=True)
    for _ in range(3):
        s.compute_coulomb(nao, eri, dm, J)
    s.synchronize()
    e0, e1 = cuda_rt.Event(), cuda_rt.Event()
    reps = 10
    e0.record(s.stream)
    for _ in range(reps):
        s.compute_coulomb(nao, eri, dm, J)
    e1.record(s.stream)
    e1.synchronize()
    ms = e0.elapsed_ms(e1) / reps
    gbs = 8.0 * n2 * n2 / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "coulomb_gemv", "nao": nao, "eri_bytes": 8 * n2 * n2, "ms": ms, "achieved_gbs": gbs,
                      "frac_of_hbm": gbs / peaks["hbm_gbs"]}), flush=True)
    K = DeviceArray((nao, nao), zero=True)
    for _ in range(3):
        s.compute_coulomb_exchange(nao, eri, dm, J, K)
    s.synchronize()
    e0.record(s.stream)
    for _ in range(reps):
        s.compute_coulomb_exchange(nao, eri, dm, J, K)
    e1.record(s.stream)
    e1.synchronize()
    ms = e0.elapsed_ms(e1) / reps
    gbs = 8.0 * n2 * n2 / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "coulomb_exchange (J and K, one ERI pass)", "nao": nao, "eri_bytes": 8 * n2 * n2, "ms": ms,
                      "achieved_gbs": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"]}), flush=True)
    eri.free()
