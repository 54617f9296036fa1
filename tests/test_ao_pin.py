"""Independent pin of the AO evaluators: tests/golden/ao_pin_table.json holds values and gradients of one s and one p
contracted shell at ten points, computed with mpmath at 50 digits directly from the textbook formulas by
tools/make_ao_pin_table.py, which imports nothing from this repository.  The CPU oracle, the numpy host statement and
(on the GPU) DFT_EvalAO must all reproduce it; the shell tables handed to them are built HERE from the JSON's raw
exponents / contraction coefficients with the normalisation formula written out, not through molgrid.sto3g_basis."""
import json
import math
import os
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table():
    t = json.load(open(os.path.join(ROOT, "tests", "golden", "ao_pin_table.json")))
    ex = np.array([float(x) for x in t["exps"]])
    cs = np.array([float(x) for x in t["coef_s"]]) * (2.0 * ex / math.pi) ** 0.75
    cp = np.array([float(x) for x in t["coef_p"]]) * (128.0 * ex ** 5 / math.pi ** 3) ** 0.25
    R = np.array([float(x) for x in t["centre"]])
    basis = SimpleNamespace(nshell=2, nao=4, shell_xyz=np.array([R, R]), shell_l=np.array([0, 1], dtype=np.int32),
                            shell_ao_off=np.array([0, 1], dtype=np.int32), shell_prim_off=np.array([0, 3], dtype=np.int32),
                            shell_nprim=np.array([3, 3], dtype=np.int32), prim_exp=np.concatenate([ex, ex]),
                            prim_coef=np.concatenate([cs, cp]), shell_atom=np.array([0, 0], dtype=np.int32))
    pts = np.array([[float(x) for x in r["point"]] for r in t["rows"]])
    val = np.array([[float(r["s"])] + [float(x) for x in r["p"]] for r in t["rows"]])               # (10, 4)
    grad = np.array([[[float(r["s_grad"][i])] + [float(x) for x in r["p_grad"][i]] for r in t["rows"]]
                     for i in range(3)])                                                               # (3, 10, 4)
    return basis, pts, val, grad


def _close(got, want, what):
    # relative 1e-13 on every entry; the absolute floor only forgives what screening (a r^2 > 60) may drop
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-26, err_msg=what)


def test_table_is_nontrivial():
    basis, pts, val, grad = _table()
    assert val.shape == (10, 4) and grad.shape == (3, 10, 4)
    assert np.all(np.abs(val[:, 0]) > 0) and np.abs(val).max() > 0.1 and np.abs(val[-1]).max() < 1e-3
    assert val[1, 2] == 0.0 and val[1, 3] == 0.0            # the point on the p_y / p_z nodal planes


def test_oracle_and_numpy_evaluators_match_the_pin(oracle):
    from quantum_compute_dft_b200 import molgrid as M
    basis, pts, val, grad = _table()
    ao, g = oracle.eval_ao(pts, basis, deriv=1)
    _close(ao, val, "oracle value"); _close(g, grad, "oracle gradient")
    ao_n, g_n = M.eval_ao_numpy(pts, basis, deriv=1)
    _close(ao_n, val, "numpy value"); _close(g_n, grad, "numpy gradient")


@pytest.mark.gpu
def test_gpu_evaluator_matches_the_pin(engine_lib):
    from quantum_compute_dft_b200.cuda_rt import DeviceArray
    from quantum_compute_dft_b200.solver import DFTSolverWrapper
    basis, pts, val, grad = _table()
    s = DFTSolverWrapper(engine_lib, "GGA")
    reps = 37                                   # several blocks of the kernel, ragged tail
    big = np.tile(pts, (reps, 1))
    for shape in (0, 1, 8, 16, 17, 32):
        s.set_option("ao_shape", shape)
        d_c = DeviceArray.from_host(big)
        d_ao, d_g = DeviceArray((big.shape[0], 4)), DeviceArray((3, big.shape[0], 4))
        s.eval_ao(d_c, basis, d_ao, d_g)
        _close(d_ao.get(), np.tile(val, (reps, 1)), f"GPU value, block shape {shape}")
        _close(d_g.get(), np.tile(grad, (1, reps, 1)), f"GPU gradient, block shape {shape}")
        d_ao0 = DeviceArray((big.shape[0], 4))
        s.eval_ao(d_c, basis, d_ao0, None)
        _close(d_ao0.get(), np.tile(val, (reps, 1)), f"GPU value only, block shape {shape}")
