"""The PRODUCT's pointwise functional header, compiled for the host, against the oracle (no GPU)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "xcfun_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-shared", "-fPIC", "-x", "c++",
                           "-I" + os.path.join(ROOT, "quantum_compute_dft_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_shim", "xcfun_host.cpp"), "-o", out])
    return ctypes.CDLL(out)


@pytest.mark.parametrize("xc,exact", [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0)])
def test_engine_functionals_match_oracle(oracle, host_shim, xc, exact):
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(1)
    n = 100000
    rho = 10 ** rng.uniform(-13, 2.5, n)
    # physical regime: |grad rho| bounded relative to rho^(4/3) (avoids the catastrophic-cancellation corner)
    g = rng.standard_normal((3, n)) * rho ** (4.0 / 3.0) * 10 ** rng.uniform(-3, 1.5, n)
    g = np.ascontiguousarray(g)
    w = rng.uniform(0, 2, n)
    sig = (g ** 2).sum(0)
    out = np.zeros((n, 5))
    host_shim.xcfun_host_eval(xc, exact, ctypes.c_long(n), rho.ctypes.data_as(dp), g[0].ctypes.data_as(dp),
                              g[1].ctypes.data_as(dp), g[2].ctypes.data_as(dp), w.ctypes.data_as(dp),
                              out.ctypes.data_as(dp))
    exc, vr, vs = oracle.functional_points(xc, rho, sig, mode=exact, gate=True)
    gate = rho >= 1e-12
    np.testing.assert_allclose(out[:, 0], w * exc, rtol=2e-11, atol=1e-300)
    np.testing.assert_allclose(out[:, 1], np.where(gate, 0.5 * w * vr, 0.0), rtol=1e-9, atol=1e-14)
    if xc:
        bref = np.where(gate, 2.0 * w * vs, 0.0) * g
        # vsigma has an exchange/correlation cancellation at tiny sigma (exact in the exact-beta mode):
        # there only the absolute size matters
        np.testing.assert_allclose(out[:, 2:].T, bref, rtol=1e-6, atol=1e-15)
        tight = sig > 1e-10
        np.testing.assert_allclose(out[tight, 2:].T, bref[:, tight], rtol=1e-8, atol=1e-13)
    assert np.all(out[~gate] == 0.0)
