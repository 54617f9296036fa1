"""In-tree build of the engine: csrc/*.cu -> quantum_compute_dft_b200/weights/dft.so.

Mirrors the reference's build line (README.md:63:
  nvcc -O3 -shared -Xcompiler -fPIC -arch=sm_80 -I./eigen_lib -I./src -lcublas ./src/dft_solver.cu -o ./weights/dft.so)
with the architecture swapped for sm_100a and without Eigen / cuBLAS.  The output name and
relative location (`weights/dft.so`) are the ones dft.py:107 loads.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "weights", "dft.so")
SOURCES = ["capi.cu", "xc_generic.cu", "xc_tma.cu", "xc_small.cu", "ao_eval.cu", "linalg.cu", "microbench.cu", "comm.cu", "fanout.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"] + os.environ.get("DFT_EXTRA_NVCC_FLAGS", "").split()


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, verbose, extra=(), tag=""):
    obj = os.path.join(OBJ, src.replace(".cu", tag + ".o"))
    spath = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), _deps_mtime()):
        return obj, ""
    cmd = ["nvcc"] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", spath, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}")
    return obj, r.stdout


def build(verbose=False, force=False, extra=(), tag=""):
    """Build weights/dft.so; `extra` nvcc flags with a non-empty `tag` build a diagnostic variant
    weights/dft<tag>.so beside it (e.g. extra=["-DDFT_PHASE_TIMING"], tag="_timing")."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    out = OUT.replace(".so", tag + ".so")
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose, extra, tag), SOURCES))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if (not os.path.exists(out)) or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    if "--timing" in sys.argv:
        print(build(verbose="-v" in sys.argv, extra=["-DDFT_PHASE_TIMING", "-DDFT_DIAGNOSTICS", "-DDFT_V_EXPERIMENTS"], tag="_timing"))
    elif "--diag" in sys.argv:
        print(build(verbose="-v" in sys.argv, extra=["-DDFT_DIAGNOSTICS", "-DDFT_V_EXPERIMENTS"], tag="_diag"))
    else:
        print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
