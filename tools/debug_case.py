#!/usr/bin/env python3
"""Debug helper: compare the generic path with the TMA path (all V tile shapes) on one random case.
Usage: python tools/debug_case.py <LDA|GGA|B3LYP> <ngrid> <nao> [seed]"""
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
variants = (("generic", {"path": 1}), ("tma", {}), ("tma-again", {}), ("vk16", {"vxc_vk": 16}), ("v64", {"vxc_shape": 64}),
            ("v128", {"vxc_shape": 128}), ("v160", {"vxc_shape": 160}), ("noskip", {"zero_skip": 0}), ("no3d", {"tma_3d": 0}))
for name, opt in variants:
    e, v, s = _run_engine(DEFAULT_LIB, fn, dm, ao, w, grad, opt)
    res[name] = (e, v)
    print(f"{name:10s} path {int(s['path'])} E = {e!r}  |V|max = {np.abs(v).max():.6e}  "
          f"dE {e - res['generic'][0]:.3e} dV {np.abs(v - res['generic'][1]).max():.3e}", flush=True)
