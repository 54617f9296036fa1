"""Single-process multi-GPU behind the unmodified ABI (csrc/fanout.cu; SURVEY.md 8e "process model").

The reference's driver keeps all arrays on one device and calls DFT_ComputeXC from one process
(dft.py:155-176, :205-208).  With option "devices" (or DFT_B200_DEVICES) that same call deals the grid to
one child engine per GPU.  "virtual_devices" runs the identical mechanism with several children per GPU, so
the deal, the resident-shard cache, the fixed-order reduction and the statistics are all exercised on a
one-GPU box; the tests that need two physical GPUs skip themselves there.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_TOL, V_TOL = 1e-8, 1e-9
XC = {"LDA": 0, "GGA": 1, "B3LYP": 2}


# ------------------------------------------------------------------------------------------ host logic (CPU)
def test_shard_points_match_the_python_deal(engine_lib):
    """DFT_ShardPoints (the C++ deal of fanout.cu) == len(solver.shard_indices) for every shard, and the shards
    cover the grid exactly once."""
    from quantum_compute_dft_b200.solver import load_library, shard_indices
    lib = load_library(engine_lib)
    for ngrid in (0, 1, 1023, 1024, 1025, 2048, 34310, 143556, 655136, 1436406):
        for n in (1, 2, 3, 4, 8):
            got = [lib.DFT_ShardPoints(ngrid, n, r) for r in range(n)]
            want = [int(shard_indices(ngrid, r, n).size) for r in range(n)]
            assert got == want, (ngrid, n)
            assert sum(got) == ngrid
    assert lib.DFT_ShardPoints(10, 0, 0) == -1 and lib.DFT_ShardPoints(10, 2, 2) == -1 and lib.DFT_ShardPoints(-1, 2, 0) == -1


# ------------------------------------------------------------------------------------------ GPU
def _case(rng, ngrid, nao):
    scale = 10 ** rng.uniform(-6, 0, (ngrid, 1))
    ao = rng.standard_normal((ngrid, nao)) * scale
    grad = rng.standard_normal((3, ngrid, nao)) * scale
    C = rng.standard_normal((nao, max(1, nao // 2))) / np.sqrt(nao)
    dm = 2.0 * C @ C.T
    w = rng.uniform(0.0, 1.0, ngrid)
    return dm, ao, w, grad


class _Run:
    def __init__(self, lib_path, functional, dm, ao, w, grad, options=()):
        from quantum_compute_dft_b200.cuda_rt import DeviceArray
        from quantum_compute_dft_b200.solver import DFTSolverWrapper
        self.ngrid, self.nao = ao.shape
        self.s = DFTSolverWrapper(lib_path, functional)
        for k, v in options:
            self.s.set_option(k, v)
        self.d_dm, self.d_ao, self.d_w = DeviceArray.from_host(dm), DeviceArray.from_host(ao), DeviceArray.from_host(w)
        self.d_g = DeviceArray.from_host(grad) if functional != "LDA" else None
        self.d_v = DeviceArray((self.nao, self.nao), zero=True)

    def __call__(self):
        e = self.s.compute_xc(self.ngrid, self.nao, self.d_dm, self.d_ao, self.d_w, self.d_v, self.d_g)
        return e, self.d_v.get()


FAN = (("virtual_devices", 3), ("devices_min_work", 0))


@pytest.mark.gpu
@pytest.mark.parametrize("functional", ["LDA", "GGA", "B3LYP"])
@pytest.mark.parametrize("ngrid,nao,nchild", [(8192, 64, 2), (10001, 33, 3), (12345, 152, 3), (7000, 7, 2), (20481, 36, 4)])
def test_fanout_equals_single_engine_and_oracle(oracle, engine_lib, functional, ngrid, nao, nchild):
    rng = np.random.default_rng(ngrid + nao)
    dm, ao, w, grad = _case(rng, ngrid, nao)
    e1, v1 = _Run(engine_lib, functional, dm, ao, w, grad)()
    fan = _Run(engine_lib, functional, dm, ao, w, grad, (("virtual_devices", nchild), ("devices_min_work", 0)))
    e, v = fan()
    assert fan.s.stat("fan_active") == 1 and fan.s.stat("devices") == nchild and fan.s.stat("fan_scatters") == 1
    assert fan.s.stat("launches") > 2
    assert abs(e - e1) <= 1e-10 * max(1.0, abs(e1))
    np.testing.assert_allclose(v, v1, rtol=0, atol=1e-11 * max(1.0, np.abs(v1).max()))
    np.testing.assert_array_equal(v, v.T)
    e_o, v_o = oracle.compute_xc(XC[functional], dm, ao, w, grad)
    assert abs(e - e_o) <= E_TOL
    np.testing.assert_allclose(0.5 * (v + v.T), oracle.sym(v_o), rtol=0, atol=V_TOL)
    # bit-reproducible: fixed-order reduction over the children (from the second call on -- after the first one every
    # engine, fanned out or not, may switch its V kernel to the dense instance: capi.cu apply_counters)
    e2, v2 = fan()
    e3, v3 = fan()
    assert e3 == e2 and abs(e2 - e) <= 1e-10 * max(1.0, abs(e))
    np.testing.assert_array_equal(v3, v2)
    assert fan.s.stat("fan_scatters") == 1          # same arrays: the resident shards were reused


@pytest.mark.gpu
def test_resident_shards_follow_the_callers_arrays(oracle, engine_lib):
    """SCF pattern: D changes every call (no re-cut); AO arrays rewritten in place are noticed (fingerprint);
    "ao_cache" 0 re-cuts every call; "devices" 1 goes back to the plain single-GPU call."""
    rng = np.random.default_rng(7)
    dm, ao, w, grad = _case(rng, 9000, 40)
    fan = _Run(engine_lib, "GGA", dm, ao, w, grad, FAN)
    fan()
    dm2 = 0.5 * dm + 0.1 * np.eye(40)
    fan.d_dm.set(dm2)
    e, v = fan()
    assert fan.s.stat("fan_scatters") == 1
    e_o, v_o = oracle.compute_xc(1, dm2, ao, w, grad)
    assert abs(e - e_o) <= E_TOL
    np.testing.assert_allclose(0.5 * (v + v.T), oracle.sym(v_o), rtol=0, atol=V_TOL)
    # new contents behind the same pointers
    dm3, ao3, w3, grad3 = _case(np.random.default_rng(8), 9000, 40)
    fan.d_ao.set(ao3); fan.d_g.set(grad3); fan.d_w.set(w3)
    e, v = fan()
    assert fan.s.stat("fan_scatters") == 2
    e_o, v_o = oracle.compute_xc(1, dm2, ao3, w3, grad3)
    assert abs(e - e_o) <= E_TOL
    np.testing.assert_allclose(0.5 * (v + v.T), oracle.sym(v_o), rtol=0, atol=V_TOL)
    # only the weights change (all of them are in the fingerprint)
    w4 = w3.copy(); w4[4321] *= 1.5
    fan.d_w.set(w4)
    e4, v4 = fan()
    assert fan.s.stat("fan_scatters") == 3
    assert abs(e4 - oracle.compute_xc(1, dm2, ao3, w4, grad3)[0]) <= E_TOL
    fan.s.set_option("ao_invalidate", 1)
    fan()
    assert fan.s.stat("fan_scatters") == 4
    fan.s.set_option("ao_cache", 0)
    fan(); fan()
    assert fan.s.stat("fan_scatters") == 6
    assert fan.s.stat("fan_resident_bytes") >= 8 * 9000 * 40 * 4
    fan.s.set_option("virtual_devices", 1)
    e1, v1 = fan()
    assert fan.s.stat("fan_active") == 0 and fan.s.stat("devices") == 1
    assert abs(e1 - e4) <= 1e-10
    np.testing.assert_allclose(v1, v4, rtol=0, atol=1e-11 * max(1.0, np.abs(v4).max()))


@pytest.mark.gpu
def test_small_builds_stay_on_the_callers_device(engine_lib):
    rng = np.random.default_rng(3)
    dm, ao, w, grad = _case(rng, 9000, 16)
    fan = _Run(engine_lib, "LDA", dm, ao, w, grad, (("virtual_devices", 2),))   # default devices_min_work
    fan()
    assert fan.s.stat("fan_active") == 0 and fan.s.stat("devices") == 2
    with pytest.raises(ValueError):
        fan.s.set_option("devices", 99)


@pytest.mark.gpu
def test_options_reach_the_children_and_raw_convention(oracle, engine_lib):
    """Engine options set before or after "devices" apply on every device: the generic path and GGA's raw
    (unsymmetrised) output convention give the same numbers fanned out as on one device."""
    rng = np.random.default_rng(11)
    dm, ao, w, grad = _case(rng, 7000, 24)
    one = _Run(engine_lib, "GGA", dm, ao, w, grad, (("raw_convention", 1), ("path", 1)))
    e1, v1 = one()
    fan = _Run(engine_lib, "GGA", dm, ao, w, grad, (("raw_convention", 1),) + FAN + (("path", 1),))
    e, v = fan()
    assert fan.s.stat("fan_active") == 1 and fan.s.stat("path") == 1
    assert np.abs(v1 - v1.T).max() > 1e-6            # really the raw convention
    assert abs(e - e1) <= 1e-10 * max(1.0, abs(e1))
    np.testing.assert_allclose(v, v1, rtol=0, atol=1e-11 * max(1.0, np.abs(v1).max()))


@pytest.mark.gpu
def test_async_entry_point_fanned_out(oracle, engine_lib):
    from quantum_compute_dft_b200.cuda_rt import DeviceArray
    rng = np.random.default_rng(12)
    dm, ao, w, grad = _case(rng, 8192, 32)
    fan = _Run(engine_lib, "B3LYP", dm, ao, w, grad, FAN)
    d_e = DeviceArray((1,), zero=True)
    fan.s.compute_xc_async(fan.ngrid, fan.nao, fan.d_dm, fan.d_ao, fan.d_w, fan.d_v, d_e, fan.d_g)
    fan.s.synchronize()
    e_o, v_o = oracle.compute_xc(2, dm, ao, w, grad)
    assert abs(d_e.get()[0] - e_o) <= E_TOL
    v = fan.d_v.get()
    np.testing.assert_allclose(0.5 * (v + v.T), oracle.sym(v_o), rtol=0, atol=V_TOL)


@pytest.mark.gpu
def test_config_molecule_fanned_out(oracle, engine_lib):
    """DHA (nao 152, TMA path, real AO sparsity) on the engine's own AO evaluator: fan-out == one device."""
    from quantum_compute_dft_b200 import workload
    hp = workload.host_problem("C4", scale=0.06)
    one = workload.make_solver(hp.functional, engine_lib)
    dp = workload.device_problem(hp, one)
    e1 = one.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
    v1 = dp.d_vxc.get()
    fan = workload.make_solver(hp.functional, engine_lib)
    for k, v in FAN:
        fan.set_option(k, v)
    for _ in range(2):   # (the second call runs the re-dealt V kernel on every child)
        e = fan.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
        v = dp.d_vxc.get()
        assert fan.stat("fan_active") == 1 and fan.stat("path") == 2
        assert abs(e - e1) <= 1e-10
        np.testing.assert_allclose(v, v1, rtol=0, atol=1e-11)
    assert fan.stat("fan_scatters") == 1
    dp.free()


def _two_gpus():
    from quantum_compute_dft_b200 import cuda_rt
    return cuda_rt.device_count() >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("functional", ["LDA", "B3LYP"])
def test_two_physical_devices(oracle, engine_lib, functional):
    """Needs two GPUs: shards cross NVLink, the reduction reads peer memory, the caller's device stays current."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices")
    import ctypes
    from quantum_compute_dft_b200 import cuda_rt
    rng = np.random.default_rng(21)
    dm, ao, w, grad = _case(rng, 50001, 152)
    cuda_rt.set_device(0)
    e1, v1 = _Run(engine_lib, functional, dm, ao, w, grad)()
    os.environ["DFT_B200_DEVICES"] = "2"          # what a user of the unmodified dft.py would set
    try:
        fan = _Run(engine_lib, functional, dm, ao, w, grad, (("devices_min_work", 0),))
    finally:
        del os.environ["DFT_B200_DEVICES"]
    assert fan.s.stat("devices") == 2
    for _ in range(3):
        e, v = fan()
    cur = ctypes.c_int(-1)
    cuda_rt.rt().cudaGetDevice(ctypes.byref(cur))
    assert cur.value == 0
    assert fan.s.stat("fan_active") == 1 and fan.s.stat("fan_scatters") == 1
    assert abs(e - e1) <= 1e-10 * max(1.0, abs(e1))
    np.testing.assert_allclose(v, v1, rtol=0, atol=1e-11 * max(1.0, np.abs(v1).max()))
    e_o, v_o = oracle.compute_xc(XC[functional], dm, ao, w, grad)
    assert abs(e - e_o) <= E_TOL
    np.testing.assert_allclose(0.5 * (v + v.T), oracle.sym(v_o), rtol=0, atol=V_TOL)
