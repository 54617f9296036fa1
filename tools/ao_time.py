"""Time DFT_EvalAO (subsystem a) on a workload's full grid: python tools/ao_time.py C5 [C4 ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_compute_dft_b200 import workload  # noqa: E402

for wl in sys.argv[1:]:
    hp = workload.host_problem(wl)
    solver = workload.make_solver(hp.functional)
    solver.set_option("timing", 1)
    dp = workload.device_problem(hp, solver)
    P = 1 if hp.functional == "LDA" else 4
    nbytes = 8.0 * dp.ngrid * dp.nao * P
    for order in (0,):
        solver.set_option("ao_input_order", order)
        for shape in (0, 16, 32):
            try:
                solver.set_option("ao_shape", shape % 100); solver.set_option("ao_vec_stores", 1 if shape >= 100 else 0)
                best = 1e9
                for _ in range(4):
                    solver.eval_ao(dp.d_coords, hp.basis, dp.d_ao, dp.d_ao_grad)
                    best = min(best, solver.stat("ao_ms"))
                print(f"{wl} {hp.name} ao_input_order={order} ao_shape={shape}: DFT_EvalAO {best:.3f} ms, {nbytes / 1e9:.2f} GB written, "
                      f"{nbytes / best / 1e9:.2f} TB/s", flush=True)
            except Exception as e:   # a shape whose staging does not fit in shared memory
                print(f"{wl} ao_shape={shape}: {e}", flush=True)
    solver.set_option("ao_input_order", 0); solver.set_option("ao_vec_stores", 0)
    solver.set_option("ao_shape", 0)
    dp.free()
