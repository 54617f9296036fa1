"""Pins the CPU oracle (and, on the GPU, the engine) to outputs of the reference's own CUDA code.

tests/golden/ref_*.npz were produced on a B200 by tools/make_reference_golden.py from
oracle/_ref/dft_ref.so (= /root/reference/src/dft_solver.cu compiled unmodified for sm_100a)."""
import os

import numpy as np
import pytest

from golden_cases import CASES, ROOT, build_case

XC = {"LDA": 0, "GGA": 1, "B3LYP": 2}


def _golden(name):
    p = os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated yet (needs one GPU run of tools/make_reference_golden.py)")
    return np.load(p)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("functional", ["LDA", "GGA", "B3LYP"])
def test_oracle_matches_reference_cuda(oracle, name, functional):
    g = _golden(name)
    dm, ao, w, grad = build_case(name)
    np.testing.assert_allclose([dm.sum(), ao.sum(), w.sum(), grad.sum()], g["input_checksum"], rtol=1e-9)
    e, v = oracle.compute_xc(XC[functional], dm, ao, w, grad, mode=oracle.COMPAT)
    assert abs(e - float(g[f"exc_{functional}"])) <= 1e-10 * max(1.0, abs(e))
    vr = g[f"vxc_raw_{functional}"]
    # raw conventions too: LDA symmetric, GGA unsymmetrised B^T Phi, B3LYP M + M^T
    np.testing.assert_allclose(v, vr, rtol=0, atol=1e-10 * max(1.0, np.abs(vr).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("functional", ["LDA", "GGA", "B3LYP"])
def test_engine_matches_reference_golden(engine_lib, name, functional):
    from test_gpu_parity import _run_engine
    g = _golden(name)
    dm, ao, w, grad = build_case(name)
    e, v, _ = _run_engine(engine_lib, functional, dm, ao, w, grad)
    vr = g[f"vxc_raw_{functional}"]
    assert abs(e - float(g[f"exc_{functional}"])) <= 1e-8
    np.testing.assert_allclose(0.5 * (v + v.T), 0.5 * (vr + vr.T), rtol=0, atol=1e-9)
