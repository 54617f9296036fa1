"""Builds the inputs of one XC workload (BASELINE.json configs C1..C5) on the current GPU.

Host side: geometry, STO-3G tables, synthetic grid and density matrix (molgrid.py).
Device side: coordinates/weights/D are uploaded, AO values and gradients are produced in
place by the engine's own DFT_EvalAO -- at C5 the AO planes are 17.3 GB, far too large to
generate with numpy and push over PCIe inside a benchmark that must finish in minutes.
"""
from dataclasses import dataclass

import numpy as np

from . import molgrid
from .cuda_rt import DeviceArray
from .solver import DFTSolverWrapper, shard_bounds, shard_indices

FUNCTIONAL_TYPE = {"LDA": 0, "GGA": 1, "B3LYP": 2}


@dataclass
class HostProblem:
    name: str
    functional: str
    mol: object
    basis: object
    coords: np.ndarray
    weights: np.ndarray
    dm: np.ndarray

    @property
    def ngrid(self):
        return self.coords.shape[0]

    @property
    def nao(self):
        return self.basis.nao


def host_problem(workload, scale=1.0, seed=0, functional=None):
    """workload: 'C1'..'C5' or a molecule name (then `functional` must be given)."""
    if workload in molgrid.WORKLOADS:
        fn, molname = molgrid.WORKLOADS[workload]
        if functional:
            fn = functional
    else:
        if not functional:
            raise ValueError("functional required when a molecule name is given")
        fn, molname = functional, workload
    mol = molgrid.load_molecule(molname)
    basis = molgrid.sto3g_basis(mol)
    coords, weights, _ = molgrid.make_grid(mol, scale=scale)
    S = molgrid.overlap_matrix(basis)
    dm = molgrid.synthetic_density(S, mol.nocc, seed=seed)
    return HostProblem(f"{fn}/{molname}", fn.upper(), mol, basis, coords, weights, dm)


@dataclass
class DeviceProblem:
    host: HostProblem
    ngrid: int          # points held by THIS rank
    ngrid_total: int
    nao: int
    d_coords: DeviceArray
    d_weights: DeviceArray
    d_dm: DeviceArray
    d_ao: DeviceArray
    d_ao_grad: object   # DeviceArray or None
    d_vxc: DeviceArray

    def free(self):
        for a in (self.d_coords, self.d_weights, self.d_dm, self.d_ao, self.d_ao_grad, self.d_vxc):
            if a is not None:
                a.free()


def device_problem(hp, solver, rank=0, nranks=1):
    """Upload this rank's share of the grid (interleaved blocks, solver.shard_indices) and evaluate its AO
    planes on the GPU."""
    idx = shard_indices(hp.ngrid, rank, nranks)
    n = int(idx.size)
    nao = hp.nao
    if nranks > 1 and n % 2 == 1 and nao % 2 == 1:
        # odd x odd would leave the TMA path: give the last point a zero-weight twin
        idx = np.concatenate([idx, idx[-1:]])
        n += 1
        w_local = hp.weights[idx].copy()
        w_local[-1] = 0.0
    else:
        w_local = hp.weights[idx]
    d_coords = DeviceArray.from_host(np.ascontiguousarray(hp.coords[idx]))
    d_w = DeviceArray.from_host(np.ascontiguousarray(w_local))
    d_dm = DeviceArray.from_host(hp.dm)
    d_ao = DeviceArray((n, nao))
    d_grad = DeviceArray((3, n, nao)) if hp.functional != "LDA" else None
    d_vxc = DeviceArray((nao, nao), zero=True)
    if n > 0:
        solver.eval_ao(d_coords, hp.basis, d_ao, d_grad)
    return DeviceProblem(hp, n, hp.ngrid, nao, d_coords, d_w, d_dm, d_ao, d_grad, d_vxc)


def make_solver(functional, lib_path=None):
    return DFTSolverWrapper(lib_path, functional) if lib_path else DFTSolverWrapper(functional_type=functional)


def algorithmic_flops(ngrid, nao):
    """dense contractions per XC build, F = 4 ngrid nao^2."""
    return 4.0 * ngrid * nao * nao


def algorithmic_bytes(ngrid, nao, functional):
    """every AO plane once, weights once, D in, V out."""
    P = 1 if functional.upper() == "LDA" else 4
    return 8.0 * ngrid * (P * nao + 1) + 16.0 * nao * nao
