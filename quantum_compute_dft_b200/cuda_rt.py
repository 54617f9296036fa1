"""Minimal ctypes binding of the CUDA runtime: device arrays for the Python harness.

The reference's driver keeps its device arrays in CuPy (dft.py:155-176) and hands
`array.data.ptr` to the C ABI.  CuPy is not installable here, and the product must
not depend on PyTorch, so the harness (tests, bench, smoke) uses this ~150-line
stand-in: `DeviceArray` mimics the three CuPy features dft.py relies on --
`.data.ptr`, `.get()` and `.set()`.  Plumbing only; no computation happens here.
"""
import ctypes
import os

import numpy as np

_CANDIDATES = [
    "libcudart.so.12",
    "/usr/local/cuda/lib64/libcudart.so.12",
    "/usr/local/cuda/lib64/libcudart.so",
]

_rt = None

H2D, D2H, D2D = 1, 2, 3


class CudaError(RuntimeError):
    pass


def rt():
    global _rt
    if _rt is not None:
        return _rt
    last = None
    for c in _CANDIDATES:
        try:
            L = ctypes.CDLL(c)
            break
        except OSError as e:  # pragma: no cover
            last = e
    else:
        raise CudaError(f"cannot load the CUDA runtime: {last}")
    L.cudaGetErrorString.restype = ctypes.c_char_p
    L.cudaGetErrorString.argtypes = [ctypes.c_int]
    L.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    L.cudaFree.argtypes = [ctypes.c_void_p]
    L.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    L.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    L.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
    L.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    L.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    L.cudaFreeHost.argtypes = [ctypes.c_void_p]
    L.cudaGetDeviceCount.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.cudaSetDevice.argtypes = [ctypes.c_int]
    L.cudaEventCreate.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    L.cudaEventRecord.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.cudaEventSynchronize.argtypes = [ctypes.c_void_p]
    L.cudaEventElapsedTime.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_void_p]
    L.cudaEventDestroy.argtypes = [ctypes.c_void_p]
    L.cudaStreamSynchronize.argtypes = [ctypes.c_void_p]
    L.cudaMemGetInfo.argtypes = [ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
    _rt = L
    return L


def check(code, what=""):
    if code != 0:
        raise CudaError(f"CUDA error {code} in {what}: {rt().cudaGetErrorString(code).decode()}")


def device_count():
    try:
        n = ctypes.c_int(0)
        if rt().cudaGetDeviceCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except CudaError:
        return 0


def set_device(i):
    check(rt().cudaSetDevice(int(i)), "cudaSetDevice")


def synchronize():
    check(rt().cudaDeviceSynchronize(), "cudaDeviceSynchronize")


def mem_info():
    f, t = ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(rt().cudaMemGetInfo(ctypes.byref(f), ctypes.byref(t)), "cudaMemGetInfo")
    return f.value, t.value


class _Data:
    __slots__ = ("ptr",)

    def __init__(self, ptr):
        self.ptr = ptr


class DeviceArray:
    """C-contiguous device array with the CuPy surface dft.py uses (.data.ptr/.get/.set)."""

    def __init__(self, shape, dtype=np.float64, zero=False):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        p = ctypes.c_void_p(0)
        check(rt().cudaMalloc(ctypes.byref(p), max(self.nbytes, 8)), f"cudaMalloc({self.nbytes})")
        self.data = _Data(p.value or 0)
        if zero:
            check(rt().cudaMemset(self.data.ptr, 0, max(self.nbytes, 8)), "cudaMemset")

    @classmethod
    def from_host(cls, a, dtype=np.float64):
        a = np.ascontiguousarray(a, dtype=dtype)
        d = cls(a.shape, dtype)
        d.set(a)
        return d

    def set(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        if a.size != self.size:
            raise ValueError(f"size mismatch {a.shape} vs {self.shape}")
        if self.nbytes:
            check(rt().cudaMemcpy(self.data.ptr, a.ctypes.data, self.nbytes, H2D), "cudaMemcpy H2D")

    def get(self):
        out = np.empty(self.shape, dtype=self.dtype)
        if self.nbytes:
            check(rt().cudaMemcpy(out.ctypes.data, self.data.ptr, self.nbytes, D2H), "cudaMemcpy D2H")
        return out

    def fill_zero(self):
        check(rt().cudaMemset(self.data.ptr, 0, max(self.nbytes, 8)), "cudaMemset")

    def free(self):
        if getattr(self, "data", None) is not None and self.data.ptr:
            rt().cudaFree(self.data.ptr)
            self.data.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host buffer exposed as a numpy array (for the end-to-end bench leg)."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        self.dtype = np.dtype(dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p(0)
        check(rt().cudaHostAlloc(ctypes.byref(p), max(n, 8), 0), "cudaHostAlloc")
        self.ptr = p.value
        buf = (ctypes.c_char * max(n, 8)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if self.ptr:
                rt().cudaFreeHost(self.ptr)
                self.ptr = 0
        except Exception:
            pass


class Event:
    def __init__(self):
        p = ctypes.c_void_p(0)
        check(rt().cudaEventCreate(ctypes.byref(p)), "cudaEventCreate")
        self.h = p.value

    def record(self, stream=0):
        check(rt().cudaEventRecord(self.h, stream), "cudaEventRecord")

    def synchronize(self):
        check(rt().cudaEventSynchronize(self.h), "cudaEventSynchronize")

    def elapsed_ms(self, later):
        ms = ctypes.c_float(0)
        check(rt().cudaEventElapsedTime(ctypes.byref(ms), self.h, later.h), "cudaEventElapsedTime")
        return ms.value

    def __del__(self):
        try:
            if self.h:
                rt().cudaEventDestroy(self.h)
        except Exception:
            pass


def memcpy_async(dst_ptr, src_ptr, nbytes, kind, stream=0):
    check(rt().cudaMemcpyAsync(dst_ptr, src_ptr, nbytes, kind, stream), "cudaMemcpyAsync")


def stream_synchronize(stream=0):
    check(rt().cudaStreamSynchronize(stream), "cudaStreamSynchronize")
