#!/usr/bin/env python3
"""Debug helper: which coefficient rows differ between the two density-kernel shapes?"""
import ctypes, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_parity import _random_case
from quantum_compute_dft_b200.cuda_rt import DeviceArray
from quantum_compute_dft_b200.solver import DFTSolverWrapper, DEFAULT_LIB

fn, ngrid, nao = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(ngrid + nao)
dm, ao, w, grad = _random_case(rng, ngrid, nao)
split = nao % 2 == 1
def run(ctas):
    s = DFTSolverWrapper(DEFAULT_LIB, fn)
    s.set_option("density_ctas_per_sm", ctas)
    s.lib.DFT_DebugRead.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64]
    d = [DeviceArray.from_host(x) for x in (dm, ao, w, grad)]
    v = DeviceArray((nao, nao), zero=True)
    e = s.compute_xc(ngrid, nao, d[0], d[1], d[2], v, d[3])
    MB = 64 if ctas == 2 else 128
    subs = [((ngrid + 1) // 2, 2, 0), (ngrid // 2, 2, 1)] if split else [(ngrid, 1, 0)]
    tot = sum(((r + MB - 1) // MB) * MB for r, _, _ in subs)
    buf = np.zeros((tot, 4))
    rc = s.lib.DFT_DebugRead(s.solver, b"coef", buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    assert rc == 0, rc
    out = np.zeros((ngrid, 4)); off = 0
    for rows, gmul, gadd in subs:
        out[gadd::gmul][:rows] = buf[off:off + rows]
        off += ((rows + MB - 1) // MB) * MB
    return e, out
e1, c1 = run(1)
e2, c2 = run(2)
print("E", e1, e2)
bad = np.where(np.abs(c1 - c2).max(axis=1) > 1e-9 * (np.abs(c1).max(axis=1) + 1e-30))[0]
print("rows differing:", bad.size, "of", ngrid)
print(bad[:64])
if bad.size:
    sub_rows = bad // 2 if split else bad
    print("block(64) idx:", np.unique(sub_rows // 64)[:40], " row-in-block:", np.unique(sub_rows % 64)[:64])
    print("parity:", np.unique(bad % 2))
    for g in bad[:6]: print(g, c1[g], c2[g])
