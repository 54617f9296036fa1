"""The C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("dft_solver.h", "dft_b200_ext.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"//[^\n]*", "", txt)
        txt = re.sub(r"#ifdef DFT_DIAGNOSTICS.*?#endif", "", txt, flags=re.S)   # diagnostic builds only
        names |= set(re.findall(r"\b(DFT_[A-Za-z0-9_]+)\s*\(", txt))
    return names


def test_every_declared_symbol_is_exported(engine_lib):
    from quantum_compute_dft_b200.solver import ABI_SYMBOLS, load_library
    declared = _declared_symbols()
    assert {"DFT_CreateSolver", "DFT_DestroySolver", "DFT_ComputeXC", "DFT_ComputeCoulomb"} <= declared
    assert declared == set(ABI_SYMBOLS)
    lib = load_library(engine_lib)
    for s in declared:
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.DFT_B200_Version()
    # diagnostics (wrong-result options, workspace dumps) are not part of the product library
    assert not hasattr(lib, "DFT_DebugRead")


def test_reference_binding_signature(engine_lib):
    """argtypes/restype are the ones dft.py:27-50 sets."""
    from quantum_compute_dft_b200.solver import load_library
    lib = load_library(engine_lib)
    assert lib.DFT_ComputeXC.restype is ctypes.c_double
    assert lib.DFT_ComputeXC.argtypes == [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_uint64] * 5
    assert lib.DFT_ComputeCoulomb.argtypes == [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_uint64] * 3
    assert lib.DFT_CreateSolver.restype is ctypes.c_void_p


def test_wrapper_error_behaviour(engine_lib, tmp_path):
    import pytest
    from quantum_compute_dft_b200.solver import DFTSolverWrapper
    with pytest.raises(FileNotFoundError):
        DFTSolverWrapper(str(tmp_path / "nope.so"), "LDA")
    with pytest.raises(ValueError):
        DFTSolverWrapper(engine_lib, "MP2")


def test_null_solver_is_a_noop(engine_lib):
    """dft_solver.cu:695,711: a null solver returns 0.0 / does nothing."""
    from quantum_compute_dft_b200.solver import load_library
    lib = load_library(engine_lib)
    assert lib.DFT_ComputeXC(None, 10, 2, 0, 0, 0, 0, 0) == 0.0
    lib.DFT_ComputeCoulomb(None, 2, 0, 0, 0)
    lib.DFT_DestroySolver(None)


def test_no_cpu_fallback_without_gpu(engine_lib):
    from quantum_compute_dft_b200 import cuda_rt
    from quantum_compute_dft_b200.solver import load_library
    if cuda_rt.device_count() > 0:
        return
    lib = load_library(engine_lib)
    assert not lib.DFT_CreateSolver(0)      # fails loudly (nullptr -> RuntimeError in the wrapper)
    assert not lib.DFT_CreateSolver(7)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "quantum_compute_dft_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for needle in ("import oracle", "from oracle", "libxc_oracle", "xc_oracle.c", "oracle/_"):
                    assert needle not in txt, (f, needle)


def test_documented_option_keys_match_the_library():
    """Every DFT_SetOption / DFT_GetStat key the library handles is documented in include/dft_b200_ext.h, and
    every documented key is handled (doc / code drift check; no GPU needed)."""
    capi = open(os.path.join(ROOT, "quantum_compute_dft_b200", "csrc", "capi.cu")).read()
    handled = set(re.findall(r'strcmp\(key, "([a-z_0-9]+)"\)', capi))
    hdr = open(os.path.join(ROOT, "include", "dft_b200_ext.h")).read()
    start = hdr.index("// ---- options / statistics")
    end = hdr.index("double DFT_GetStat")
    documented = set(re.findall(r'"([a-z_0-9]+)"', hdr[start:end]))
    assert handled - documented == set(), sorted(handled - documented)
    assert documented - handled == set(), sorted(documented - handled)
