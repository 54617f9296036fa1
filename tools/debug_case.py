#!/usr/bin/env python3
"""Debug helper: compare generic / TMA (both density shapes) on one random case."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_parity import _random_case, _run_engine
from quantum_compute_dft_b200.solver import DEFAULT_LIB

fn, ngrid, nao = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else ngrid + nao
rng = np.random.default_rng(seed)
dm, ao, w, grad = _random_case(rng, ngrid, nao)
res = {}
for name, opt in (("generic", {"path": 1}), ("tma1", {"density_ctas_per_sm": 1}), ("tma2", {"density_ctas_per_sm": 2}),
                  ("tma2b", {"density_ctas_per_sm": 2}), ("tma2np", {"density_ctas_per_sm": 2, "l2_prefetch": 0})):
    e, v, s = _run_engine(DEFAULT_LIB, fn, dm, ao, w, grad, opt)
    res[name] = (e, v)
    print(f"{name:8s} path {int(s['path'])} E = {e!r}  |V|max = {np.abs(v).max():.6e}")
for k in ("tma1", "tma2", "tma2b", "tma2np"):
    print(k, "dE", res[k][0] - res["generic"][0], "dV", np.abs(res[k][1] - res["generic"][1]).max())
