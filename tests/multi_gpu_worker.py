"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), NCCL all-reduce inside DFT_ComputeXC.

Checks ON HARDWARE that the all-reduced V_xc / E_xc of a grid-sharded build equal the 1-GPU build of the same
inputs (and the reference's CUDA where oracle/_ref/dft_ref.so travelled), that every rank holds the same matrix,
and that a rank with bad arguments makes EVERY rank return NaN instead of leaving the others in the collective.
Prints one JSON line per case on rank 0; exit code 0 only if everything held.
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from quantum_compute_dft_b200 import cuda_rt, workload
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    cuda_rt.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("gloo", init_method="env://")
    cases = [("C1", 1.0), ("C2", 0.25), ("C4", 0.05), ("C5", 0.02)]
    if len(sys.argv) > 1:
        cases = [(c.split(":")[0], float(c.split(":")[1])) for c in sys.argv[1:]]
    ok_all = True
    for wl, scale in cases:
        hp = workload.host_problem(wl, scale=scale)
        solver = workload.make_solver(hp.functional)
        ids = [solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        solver.comm_init(rank, world, ids[0])
        assert solver.stat("nranks") == world
        dp = workload.device_problem(hp, solver, rank, world)
        e = solver.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
        v = dp.d_vxc.get()
        # every rank holds the same global matrix and energy
        t = torch.from_numpy(np.concatenate([v.ravel(), [e]]))
        t0 = t.clone()
        dist.broadcast(t0, src=0)
        td = torch.tensor([float((t - t0).abs().max())], dtype=torch.float64)
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        rec = {"case": f"{wl} x{scale}", "ranks": world, "ngrid": hp.ngrid, "nao": hp.nao,
               "ranks_max_abs_diff": float(td[0])}
        ok = float(td[0]) == 0.0
        if rank == 0:
            s1 = workload.make_solver(hp.functional)
            dp1 = workload.device_problem(hp, s1)
            e1 = s1.compute_xc(dp1.ngrid, dp1.nao, dp1.d_dm, dp1.d_ao, dp1.d_weights, dp1.d_vxc, dp1.d_ao_grad)
            rec["vs_single_gpu"] = bench.parity_record(e, v, e1, dp1.d_vxc.get(), "this engine, 1 GPU")
            ok = ok and rec["vs_single_gpu"]["ok"]
            ref = bench.load_reference_lib()
            if ref is not None:
                d_vref = cuda_rt.DeviceArray((hp.nao, hp.nao), zero=True)
                e_r, v_r = bench.reference_xc(ref, hp.functional, dp1.ngrid, hp.nao, dp1.d_dm, dp1.d_ao, dp1.d_ao_grad,
                                              dp1.d_weights, d_vref)
                rec["vs_reference_cuda"] = bench.parity_record(e, v, e_r, v_r, "reference CUDA, 1 GPU")
                ok = ok and rec["vs_reference_cuda"]["ok"]
            dp1.free()
        # failure on ONE rank (null AO pointer): every rank must come back with NaN, nobody may hang
        class _Null:
            class data:
                ptr = 0
        bad = rank == world - 1
        e_bad = solver.compute_xc(dp.ngrid, dp.nao, dp.d_dm, _Null if bad else dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
        tn = torch.tensor([1.0 if math.isnan(e_bad) else 0.0], dtype=torch.float64)
        dist.all_reduce(tn, op=dist.ReduceOp.MIN)
        rec["one_rank_failed_all_nan"] = bool(tn[0] == 1.0)
        # ... and the communicator still works afterwards
        e_again = solver.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
        rec["recovers"] = bool(e_again == e)
        ok = ok and rec["one_rank_failed_all_nan"] and rec["recovers"]
        tk = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64)
        dist.all_reduce(tk, op=dist.ReduceOp.MIN)
        rec["ok"] = bool(tk[0] == 1.0)
        ok_all = ok_all and rec["ok"]
        if rank == 0:
            print(json.dumps(rec), flush=True)
        solver.comm_destroy()
        dp.free()
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
