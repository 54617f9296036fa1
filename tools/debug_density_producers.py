import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from quantum_compute_dft_b200 import cuda_rt
from quantum_compute_dft_b200.cuda_rt import DeviceArray
from quantum_compute_dft_b200.solver import DFTSolverWrapper, DEFAULT_LIB
fn, ngrid, nao, prod, dyn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
lib = DEFAULT_LIB.replace(".so", "_dbg.so")
cuda_rt.set_device(0)
rng = np.random.default_rng(1)
ao = rng.standard_normal((ngrid, nao)); grad = rng.standard_normal((3, ngrid, nao))
C = rng.standard_normal((nao, max(1, nao // 2))) / np.sqrt(nao); dm = 2.0 * C @ C.T
w = rng.uniform(0, 1, ngrid)
s = DFTSolverWrapper(lib, fn)
s.set_option("path", 2); s.set_option("density_producers", prod); s.set_option("dyn_sched", dyn)
d = [DeviceArray.from_host(x) for x in (dm, ao, w, grad)]
d_v = DeviceArray((nao, nao), zero=True)
t0 = time.time()
for i in range(3):
    e = s.compute_xc(ngrid, nao, d[0], d[1], d[2], d_v, d[3] if fn != "LDA" else None)
print(fn, ngrid, nao, "producers", prod, "dyn", dyn, "E", e, "%.1f ms" % ((time.time() - t0) * 1e3), flush=True)
