"""Grid-sharded multi-rank logic on the CPU (gloo, world_size 2): every rank integrates its slice of
the grid (here with the oracle standing in for the GPU engine) and the nao x nao partial V_xc plus the
scalar E_xc are all-reduced -- the same partition (solver.shard_indices: interleaved blocks), packing ([V | E]) and reduction
the engine performs with NCCL on the GPUs (csrc/comm.cu)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, functional, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from quantum_compute_dft_b200 import workload
    from quantum_compute_dft_b200.solver import shard_indices
    hp = workload.host_problem("H2O", scale=0.1, functional=functional)
    idx = shard_indices(hp.ngrid, rank, world, block=512)
    ao, grad = O.eval_ao(hp.coords[idx], hp.basis, deriv=1)
    e, v = O.compute_xc(workload.FUNCTIONAL_TYPE[functional], hp.dm, ao, hp.weights[idx], grad)
    packed = torch.from_numpy(np.concatenate([O.sym(v).ravel(), [e]]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, f"packed_{functional}.npy"), packed.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("functional", ["LDA", "B3LYP"])
def test_two_rank_grid_sharding_sums_to_whole(oracle, tmp_path, functional):
    import torch.multiprocessing as mp
    from quantum_compute_dft_b200 import workload
    port = _free_port()
    mp.spawn(_worker, args=(2, port, functional, str(tmp_path)), nprocs=2, join=True)
    packed = np.load(tmp_path / f"packed_{functional}.npy")
    hp = workload.host_problem("H2O", scale=0.1, functional=functional)
    ao, grad = oracle.eval_ao(hp.coords, hp.basis, deriv=1)
    e, v = oracle.compute_xc(workload.FUNCTIONAL_TYPE[functional], hp.dm, ao, hp.weights, grad)
    nao = hp.nao
    assert abs(packed[-1] - e) < 1e-10
    np.testing.assert_allclose(packed[:-1].reshape(nao, nao), oracle.sym(v), rtol=0, atol=1e-11)


def test_shard_indices_partition_the_grid():
    from quantum_compute_dft_b200.solver import shard_indices
    for ngrid in (1, 2, 8191, 8192, 8193, 100003, 1436406):
        for n in (1, 2, 3, 4, 8):
            parts = [shard_indices(ngrid, r, n) for r in range(n)]
            allidx = np.sort(np.concatenate(parts))
            np.testing.assert_array_equal(allidx, np.arange(ngrid))
            sizes = [p.size for p in parts]
            assert max(sizes) - min(sizes) <= 8192
