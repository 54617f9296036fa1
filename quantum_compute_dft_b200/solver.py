"""Host-side mirror of the reference's operator interface for the XC path.

`DFTSolverWrapper` has the same name, constructor arguments, methods, argument
meaning and error behaviour as the class in the reference's driver
(dft.py:15-95): it loads the shared library with ctypes, binds the four C symbols
with the very same argtypes/restype (dft.py:27-50), maps 'LDA'/'GGA'/'B3LYP' to the
solver type ints 0/1/2 (dft.py:52-59), raises FileNotFoundError / ValueError /
RuntimeError in the same situations (dft.py:21-22,59,62-63), and takes device
arrays that expose `.data.ptr` (CuPy there, cuda_rt.DeviceArray here).

The additive methods (eval_ao, comm_init, set_option, stat, ...) bind the entry
points declared in include/dft_b200_ext.h.  There is no CPU fallback: without the
built library or without a GPU the constructor raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "weights", "dft.so")

# every symbol include/dft_solver.h and include/dft_b200_ext.h declare with C linkage
ABI_SYMBOLS = [
    "DFT_CreateSolver", "DFT_DestroySolver", "DFT_ComputeXC", "DFT_ComputeCoulomb",
    "DFT_EvalAO", "DFT_CommGetUniqueId", "DFT_CommInit", "DFT_CommDestroy",
    "DFT_SetOption", "DFT_GetStat", "DFT_ComputeXCAsync", "DFT_StreamSynchronize", "DFT_GetStream",
    "DFT_MicrobenchDMMA", "DFT_MicrobenchDFMA", "DFT_MicrobenchDMMAWarps", "DFT_B200_Version",
    "DFT_ComputeCoulombExchange", "DFT_BuildFock", "DFT_SCFEnergies", "DFT_ShardPoints",
]

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int)


def load_library(lib_path=DEFAULT_LIB):
    """ctypes.CDLL with all argtypes bound.  Raises FileNotFoundError if the .so is not built."""
    if not os.path.exists(lib_path):
        raise FileNotFoundError(f"Shared library not found at: {lib_path}")
    lib = ctypes.CDLL(os.path.abspath(lib_path))
    # -- the reference ABI, bound exactly as dft.py:27-50 does
    lib.DFT_CreateSolver.argtypes = [ctypes.c_int]
    lib.DFT_CreateSolver.restype = ctypes.c_void_p
    lib.DFT_DestroySolver.argtypes = [ctypes.c_void_p]
    lib.DFT_DestroySolver.restype = None
    lib.DFT_ComputeXC.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                  ctypes.c_uint64, ctypes.c_uint64]
    lib.DFT_ComputeXC.restype = ctypes.c_double
    lib.DFT_ComputeCoulomb.argtypes = [ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
    lib.DFT_ComputeCoulomb.restype = None
    # -- additive entry points
    lib.DFT_EvalAO.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int,
                               _c_dp, _c_ip, _c_ip, _c_ip, _c_ip, ctypes.c_int, _c_dp, _c_dp,
                               ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64]
    lib.DFT_EvalAO.restype = ctypes.c_int
    lib.DFT_CommGetUniqueId.argtypes = [ctypes.c_void_p]
    lib.DFT_CommGetUniqueId.restype = ctypes.c_int
    lib.DFT_CommInit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.DFT_CommInit.restype = ctypes.c_int
    lib.DFT_CommDestroy.argtypes = [ctypes.c_void_p]
    lib.DFT_CommDestroy.restype = ctypes.c_int
    lib.DFT_SetOption.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
    lib.DFT_SetOption.restype = ctypes.c_int
    lib.DFT_GetStat.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.DFT_GetStat.restype = ctypes.c_double
    lib.DFT_ComputeXCAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_uint64] * 6
    lib.DFT_ComputeXCAsync.restype = ctypes.c_int
    lib.DFT_StreamSynchronize.argtypes = [ctypes.c_void_p]
    lib.DFT_StreamSynchronize.restype = ctypes.c_int
    lib.DFT_GetStream.argtypes = [ctypes.c_void_p]
    lib.DFT_GetStream.restype = ctypes.c_uint64
    lib.DFT_MicrobenchDMMA.argtypes = [ctypes.c_int]
    lib.DFT_MicrobenchDMMA.restype = ctypes.c_double
    lib.DFT_MicrobenchDFMA.argtypes = [ctypes.c_int]
    lib.DFT_MicrobenchDFMA.restype = ctypes.c_double
    lib.DFT_MicrobenchDMMAWarps.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.DFT_MicrobenchDMMAWarps.restype = ctypes.c_double
    lib.DFT_B200_Version.restype = ctypes.c_char_p
    lib.DFT_ShardPoints.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.DFT_ShardPoints.restype = ctypes.c_int
    return lib


class DFTSolverWrapper:
    TYPE_LDA = 0
    TYPE_GGA = 1
    TYPE_B3LYP = 2

    def __init__(self, lib_path=DEFAULT_LIB, functional_type='lda'):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(f"Shared library not found at: {lib_path}")
        self.lib = load_library(lib_path)
        self.functional_type = functional_type.upper()
        if self.functional_type == 'LDA':
            c_type = self.TYPE_LDA
        elif self.functional_type == 'GGA':
            c_type = self.TYPE_GGA
        elif self.functional_type == 'B3LYP':
            c_type = self.TYPE_B3LYP
        else:
            raise ValueError(f"Unsupported functional type: {self.functional_type}")
        self.solver = self.lib.DFT_CreateSolver(c_type)
        if not self.solver:
            raise RuntimeError("Failed to create C++ DFT Solver instance.")

    def __del__(self):
        if hasattr(self, 'lib') and hasattr(self, 'solver') and self.solver:
            self.lib.DFT_DestroySolver(self.solver)
            self.solver = None

    # ---- the reference's two operations (dft.py:69-95) --------------------------------------
    def compute_xc(self, ngrid, nao, d_dm, d_ao, d_weights, d_vxc, d_ao_grad=None):
        ptr_grad = d_ao_grad.data.ptr if d_ao_grad is not None else 0
        return self.lib.DFT_ComputeXC(self.solver, ngrid, nao,
                                      ctypes.c_uint64(d_dm.data.ptr), ctypes.c_uint64(d_ao.data.ptr),
                                      ctypes.c_uint64(ptr_grad), ctypes.c_uint64(d_weights.data.ptr),
                                      ctypes.c_uint64(d_vxc.data.ptr))

    def compute_coulomb(self, nao, d_eri, d_dm, d_J):
        self.lib.DFT_ComputeCoulomb(self.solver, nao, ctypes.c_uint64(d_eri.data.ptr),
                                    ctypes.c_uint64(d_dm.data.ptr), ctypes.c_uint64(d_J.data.ptr))

    # ---- additive -------------------------------------------------------------------------
    def build_fock(self, nao, d_hcore, d_J, d_vxc, d_K, c_hf, d_F):
        """F = Hcore + J + 1/2 (V + V^T) - 1/2 c_hf K on the device (dft.py:212,221-223)."""
        self.lib.DFT_BuildFock.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_uint64] * 4 + [ctypes.c_double, ctypes.c_uint64]
        self.lib.DFT_BuildFock.restype = ctypes.c_int
        rc = self.lib.DFT_BuildFock(self.solver, nao, d_hcore.data.ptr, d_J.data.ptr, d_vxc.data.ptr,
                                    d_K.data.ptr if d_K is not None else 0, float(c_hf), d_F.data.ptr)
        if rc != 0:
            raise RuntimeError(f"DFT_BuildFock failed with code {rc}")

    def scf_energies(self, nao, d_dm, d_hcore, d_J, d_K, c_hf):
        """(E_one, E_coul, E_hf) = (sum D o H, 1/2 sum D o J, -1/4 c_hf sum D o K), dft.py:230-236."""
        self.lib.DFT_SCFEnergies.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_uint64] * 4 + [ctypes.c_double, _c_dp]
        self.lib.DFT_SCFEnergies.restype = ctypes.c_int
        out = (ctypes.c_double * 3)()
        rc = self.lib.DFT_SCFEnergies(self.solver, nao, d_dm.data.ptr, d_hcore.data.ptr, d_J.data.ptr,
                                      d_K.data.ptr if d_K is not None else 0, float(c_hf), out)
        if rc != 0:
            raise RuntimeError(f"DFT_SCFEnergies failed with code {rc}")
        return out[0], out[1], out[2]

    def compute_coulomb_exchange(self, nao, d_eri, d_dm, d_J, d_K):
        """J and K = einsum('ijkl,jl->ik', eri, dm) (dft.py:218) in one pass over the ERI."""
        self.lib.DFT_ComputeCoulombExchange.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_uint64] * 4
        self.lib.DFT_ComputeCoulombExchange.restype = ctypes.c_int
        rc = self.lib.DFT_ComputeCoulombExchange(self.solver, nao, d_eri.data.ptr, d_dm.data.ptr, d_J.data.ptr, d_K.data.ptr)
        if rc != 0:
            raise RuntimeError(f"DFT_ComputeCoulombExchange failed with code {rc}")

    def eval_ao(self, d_coords, basis, d_ao, d_ao_grad=None, exp_cutoff=0.0):
        """GPU replacement of numint.eval_ao (grid.py:30,38).  basis: molgrid.Basis."""
        ngrid = d_coords.shape[0]
        xyz = np.ascontiguousarray(basis.shell_xyz, dtype=np.float64)
        l = np.ascontiguousarray(basis.shell_l, dtype=np.int32)
        ao_off = np.ascontiguousarray(basis.shell_ao_off, dtype=np.int32)
        p_off = np.ascontiguousarray(basis.shell_prim_off, dtype=np.int32)
        npr = np.ascontiguousarray(basis.shell_nprim, dtype=np.int32)
        ex = np.ascontiguousarray(basis.prim_exp, dtype=np.float64)
        co = np.ascontiguousarray(basis.prim_coef, dtype=np.float64)
        rc = self.lib.DFT_EvalAO(self.solver, ngrid, d_coords.data.ptr, int(basis.nshell),
                                 xyz.ctypes.data_as(_c_dp), l.ctypes.data_as(_c_ip), ao_off.ctypes.data_as(_c_ip),
                                 p_off.ctypes.data_as(_c_ip), npr.ctypes.data_as(_c_ip), int(ex.size),
                                 ex.ctypes.data_as(_c_dp), co.ctypes.data_as(_c_dp), int(basis.nao),
                                 1 if d_ao_grad is not None else 0, float(exp_cutoff), d_ao.data.ptr,
                                 d_ao_grad.data.ptr if d_ao_grad is not None else 0)
        if rc != 0:
            raise RuntimeError(f"DFT_EvalAO failed with code {rc}")

    def compute_xc_async(self, ngrid, nao, d_dm, d_ao, d_weights, d_vxc, d_exc, d_ao_grad=None):
        ptr_grad = d_ao_grad.data.ptr if d_ao_grad is not None else 0
        rc = self.lib.DFT_ComputeXCAsync(self.solver, ngrid, nao, d_dm.data.ptr, d_ao.data.ptr, ptr_grad,
                                         d_weights.data.ptr, d_vxc.data.ptr, d_exc.data.ptr)
        if rc != 0:
            raise RuntimeError(f"DFT_ComputeXCAsync failed with code {rc}")

    def synchronize(self):
        self.lib.DFT_StreamSynchronize(self.solver)

    @property
    def stream(self):
        return int(self.lib.DFT_GetStream(self.solver))

    def set_option(self, key, value):
        rc = self.lib.DFT_SetOption(self.solver, key.encode(), float(value))
        if rc != 0:
            raise ValueError(f"DFT_SetOption({key!r}, {value}) -> {rc}")

    def stat(self, key):
        return float(self.lib.DFT_GetStat(self.solver, key.encode()))

    def comm_unique_id(self):
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.DFT_CommGetUniqueId(buf)
        if rc != 0:
            raise RuntimeError(f"DFT_CommGetUniqueId -> {rc}")
        return bytes(buf.raw)

    def comm_init(self, rank, nranks, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        rc = self.lib.DFT_CommInit(self.solver, int(rank), int(nranks), buf)
        if rc != 0:
            raise RuntimeError(f"DFT_CommInit -> {rc}")

    def comm_destroy(self):
        self.lib.DFT_CommDestroy(self.solver)


def shard_indices(ngrid, rank, nranks, block=1024):
    """Grid points of `rank` when the grid is dealt to the ranks in INTERLEAVED blocks of `block` points
    (block b goes to rank b mod nranks).  Grid points are independent, so any partition is valid (SURVEY.md
    8e); with AO screening the cost of a point depends on how many atoms are near it, and a contiguous
    range (one end of the molecule) would not be representative: interleaving keeps the ranks balanced.
    `block` is even, so every shard keeps an even number of leading points per block (16-byte aligned AO
    rows for odd nao).  Returns a sorted int64 index array."""
    if nranks <= 1:
        return np.arange(ngrid, dtype=np.int64)
    nblk = (ngrid + block - 1) // block
    mine = np.arange(rank, nblk, nranks, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < ngrid]


def shard_bounds(ngrid, rank, nranks, align=2):
    """Contiguous grid-point range of `rank`: [r*ngrid/N, (r+1)*ngrid/N) rounded to
    `align` points so every shard starts on a 16-byte boundary of the AO rows."""
    def cut(r):
        c = (ngrid * r) // nranks
        c -= c % align
        return ngrid if r == nranks else c
    return cut(rank), cut(rank + 1)
