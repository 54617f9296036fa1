#!/usr/bin/env python3
"""Runs the reference's OWN CUDA implementation (oracle/_ref/dft_ref.so = /root/reference/src/dft_solver.cu
compiled unmodified for sm_100a by oracle/Makefile) on a B200 and records its outputs as golden vectors.

Run on the GPU box:   python tools/make_reference_golden.py gpurun_out/golden
then copy gpurun_out/golden/ref_*.npz into tests/golden/.  Inputs are regenerated from a recipe by
tests/test_golden_reference.py (tests/golden_cases.py), so only outputs (+ input checksums) are stored.
The CPU test-suite then pins the oracle to these vectors -- i.e. to what the reference itself computes.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_cases import CASES, build_case  # noqa: E402
from quantum_compute_dft_b200.cuda_rt import DeviceArray  # noqa: E402


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "dft_ref.so"))
    lib.DFT_CreateSolver.argtypes = [ctypes.c_int]; lib.DFT_CreateSolver.restype = ctypes.c_void_p
    lib.DFT_DestroySolver.argtypes = [ctypes.c_void_p]
    lib.DFT_ComputeXC.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_uint64] * 5
    lib.DFT_ComputeXC.restype = ctypes.c_double
    for name in CASES:
        dm, ao, w, grad = build_case(name)
        ngrid, nao = ao.shape
        out = {"input_checksum": np.array([dm.sum(), ao.sum(), w.sum(), grad.sum()])}
        d_dm, d_ao, d_w, d_g = (DeviceArray.from_host(x) for x in (dm, ao, w, grad))
        for t, fn in enumerate(("LDA", "GGA", "B3LYP")):
            s = lib.DFT_CreateSolver(t)
            d_v = DeviceArray((nao, nao), zero=True)
            e = lib.DFT_ComputeXC(s, ngrid, nao, d_dm.data.ptr, d_ao.data.ptr, d_g.data.ptr if t else 0,
                                  d_w.data.ptr, d_v.data.ptr)
            out[f"exc_{fn}"] = np.array(e)
            out[f"vxc_raw_{fn}"] = d_v.get()
            lib.DFT_DestroySolver(s)
            print(name, fn, "E_xc = %.12f" % e)
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), **out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
