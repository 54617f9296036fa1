"""First thing a GPU session runs (under `timeout 120`): every kernel family of the engine once, on small inputs, through
the C ABI.  Exit code 0 only if all of them finish with the right answer; a deadlocked kernel is killed by the timeout
before it can eat the session."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from quantum_compute_dft_b200 import cuda_rt  # noqa: E402
from quantum_compute_dft_b200.cuda_rt import DeviceArray  # noqa: E402
from quantum_compute_dft_b200.solver import DFTSolverWrapper  # noqa: E402


def main():
    cuda_rt.set_device(0)
    rng = np.random.default_rng(1)
    ok = True
    for fn, xc in (("LDA", 0), ("B3LYP", 2)):
        for (ngrid, nao), opts in (((3000, 36), {}), ((3000, 36), {"path": 2}), ((4000, 200), {}),
                                   ((4000, 377), {"vxc_skip": 1, "vxc_skip_mode": 4}),
                                   ((4000, 377), {"vxc_skip": 1, "vxc_skip_mode": 1}), ((4000, 377), {"vxc_skip": 0}),
                                   ((4000, 377), {"vxc_skip": 1, "vxc_skip_mode": 2}), ((20000, 377), {"vxc_skip": 1, "vxc_skip_mode": 2}),
                                   ((4000, 377), {"density_wide": 1}), ((4001, 200), {"density_wide": 1}), ((3000, 36), {"path": 2, "density_wide": 1}),
                                   ((20000, 152), {"density_wide": 1, "dyn_sched": 0}), ((9000, 377), {"density_wide": 1, "density_unit": 1}),
                                   ((3001, 77), {"path": 1})):
            scale = 10 ** rng.uniform(-6, 0, (ngrid, 1))
            ao = rng.standard_normal((ngrid, nao)) * scale
            ao[:, nao // 3: nao // 2] = 0.0
            grad = rng.standard_normal((3, ngrid, nao)) * scale
            grad[:, :, nao // 3: nao // 2] = 0.0
            C = rng.standard_normal((nao, max(1, nao // 2))) / np.sqrt(nao)
            dm = 2.0 * C @ C.T
            w = rng.uniform(0.0, 1.0, ngrid)
            s = DFTSolverWrapper(functional_type=fn)
            for k, v in opts.items():
                s.set_option(k, v)
            d = [DeviceArray.from_host(x) for x in (dm, ao, w, grad)]
            d_v = DeviceArray((nao, nao), zero=True)
            t0 = time.time()
            e = s.compute_xc(ngrid, nao, d[0], d[1], d[2], d_v, d[3] if fn != "LDA" else None)
            e = s.compute_xc(ngrid, nao, d[0], d[1], d[2], d_v, d[3] if fn != "LDA" else None)
            dt = time.time() - t0
            e_o, v_o = O.compute_xc(xc, dm, ao, w, grad)
            v = d_v.get()
            good = abs(e - e_o) < 1e-8 and np.max(np.abs(0.5 * (v + v.T) - O.sym(v_o))) < 1e-9
            ok = ok and good
            print(f"{fn} {ngrid}x{nao} {opts}: path {int(s.stat('path'))} {'ok' if good else 'WRONG'} ({dt * 1e3:.1f} ms for 2 calls)", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
