"""ctypes binding of oracle/_build/libxc_oracle.so (CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/xc_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libxc_oracle.so")

LDA, GGA, B3LYP = 0, 1, 2
COMPAT, EXACT = 0, 1
AO_EXP_CUTOFF = 60.0

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the C restatement (gcc + OpenMP).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "xc_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_functional_points.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                               _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]
        L.oracle_functional_points.restype = None
        L.oracle_density.argtypes = [ctypes.c_long, ctypes.c_int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]
        L.oracle_density.restype = None
        L.oracle_compute_xc.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_int,
                                        _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]
        L.oracle_compute_xc.restype = ctypes.c_double
        L.oracle_coulomb.argtypes = [ctypes.c_int, _c_dp, _c_dp, _c_dp]
        L.oracle_coulomb.restype = None
        L.oracle_eval_ao.argtypes = [ctypes.c_long, _c_dp, ctypes.c_int, _c_dp, _c_ip, _c_ip, _c_ip, _c_ip,
                                     _c_dp, _c_dp, ctypes.c_int, ctypes.c_double, ctypes.c_int, _c_dp, _c_dp]
        L.oracle_eval_ao.restype = None
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def functional_points(xc_type, rho, sigma=None, mode=COMPAT, gate=False):
    """(rho*eps, vrho, vsigma) exactly as the reference device routines return them."""
    rho = _f64(rho).ravel()
    n = rho.size
    sigma = np.zeros(n) if sigma is None else _f64(sigma).ravel()
    exc, vr, vs = np.empty(n), np.empty(n), np.empty(n)
    lib().oracle_functional_points(xc_type, mode, int(gate), n, _dp(rho), _dp(sigma), _dp(exc), _dp(vr), _dp(vs))
    return exc, vr, vs


def density(dm, ao, ao_grad=None):
    ao = _f64(ao)
    ngrid, nao = ao.shape
    dm = _f64(dm)
    rho = np.zeros(ngrid)
    grad = np.zeros((ngrid, 3))
    sigma = np.zeros(ngrid)
    g = _f64(ao_grad) if ao_grad is not None else None
    lib().oracle_density(ngrid, nao, _dp(dm), _dp(ao), _dp(g), _dp(rho), _dp(grad), _dp(sigma))
    return rho, grad, sigma


def compute_xc(xc_type, dm, ao, weights, ao_grad=None, mode=COMPAT, want_density=False):
    """Reference-convention (E_xc, raw V_xc).  Parity is on sym(V) = (V + V.T)/2 (dft.py:212)."""
    ao = _f64(ao)
    ngrid, nao = ao.shape
    dm = _f64(dm)
    w = _f64(weights)
    g = _f64(ao_grad) if (ao_grad is not None and xc_type != LDA) else None
    if xc_type != LDA and g is None:
        raise ValueError("GGA/B3LYP need ao_grad (3,ngrid,nao)")
    v = np.zeros((nao, nao))
    rho = np.zeros(ngrid) if want_density else None
    sig = np.zeros(ngrid) if want_density else None
    e = lib().oracle_compute_xc(xc_type, mode, ngrid, nao, _dp(dm), _dp(ao), _dp(g), _dp(w), _dp(v),
                                _dp(rho), _dp(sig))
    if want_density:
        return e, v, rho, sig
    return e, v


def sym(v):
    return 0.5 * (v + v.T)


def exchange(eri, dm):
    """K_ik = sum_jl (ij|kl) D_jl: the reference driver's cupy.einsum('ijkl,jl->ik', eri, dm), dft.py:218."""
    dm = _f64(dm)
    nao = dm.shape[0]
    return np.einsum("ijkl,jl->ik", _f64(eri).reshape(nao, nao, nao, nao), dm)


def coulomb(eri, dm):
    dm = _f64(dm)
    nao = dm.shape[0]
    eri = _f64(eri).reshape(nao * nao, nao * nao)
    J = np.zeros((nao, nao))
    lib().oracle_coulomb(nao, _dp(eri), _dp(dm), _dp(J))
    return J


def eval_ao(coords, basis, deriv=0, exp_cutoff=AO_EXP_CUTOFF):
    """basis: any object with the flat shell tables of quantum_compute_dft_b200.molgrid.Basis."""
    coords = _f64(coords)
    ngrid = coords.shape[0]
    nao = int(basis.nao)
    ao = np.zeros((ngrid, nao))
    grad = np.zeros((3, ngrid, nao)) if deriv else None
    lib().oracle_eval_ao(ngrid, _dp(coords), int(basis.nshell), _dp(_f64(basis.shell_xyz)),
                         _ip(np.ascontiguousarray(basis.shell_l, dtype=np.int32)),
                         _ip(np.ascontiguousarray(basis.shell_ao_off, dtype=np.int32)),
                         _ip(np.ascontiguousarray(basis.shell_prim_off, dtype=np.int32)),
                         _ip(np.ascontiguousarray(basis.shell_nprim, dtype=np.int32)),
                         _dp(_f64(basis.prim_exp)), _dp(_f64(basis.prim_coef)), nao,
                         float(exp_cutoff), int(deriv), _dp(ao), _dp(grad))
    return (ao, grad) if deriv else ao


def num_threads():
    return int(lib().oracle_num_threads())
