"""Minimal CuPy stand-in for the reference's driver (dft.py:2,155-178,200-221): see tests/shims/README.md."""
import numpy as np

from quantum_compute_dft_b200 import cuda_rt as _rt

float64 = np.float64


class ndarray:
    """C-contiguous float64 device array: .data.ptr / .set / .get / .reshape / .shape, like cupy.ndarray."""

    def __init__(self, base, shape):
        self._base = base            # cuda_rt.DeviceArray that owns the memory (shared by reshaped views)
        self.shape = tuple(int(s) for s in shape)
        self.data = base.data
        self.dtype = np.dtype(np.float64)

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if int(np.prod(shape)) != self.size:
            raise ValueError(f"cannot reshape {self.shape} into {shape}")
        return ndarray(self._base, shape)

    def set(self, a):
        self._base.set(np.ascontiguousarray(a, dtype=np.float64).reshape(self._base.shape))

    def get(self):
        return self._base.get().reshape(self.shape)


def asarray(a, dtype=None, order="C"):
    if isinstance(a, ndarray):
        return a
    a = np.ascontiguousarray(a, dtype=np.float64)
    return ndarray(_rt.DeviceArray.from_host(a), a.shape)


def zeros(shape, dtype=None, order="C"):
    shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
    return ndarray(_rt.DeviceArray(shape, zero=True), shape)


def einsum(subscripts, *operands):
    """Evaluated on the host (numpy) from downloaded copies; returns a device array."""
    host = [o.get() if isinstance(o, ndarray) else np.asarray(o) for o in operands]
    return asarray(np.einsum(subscripts, *host))


class _Stream:
    def synchronize(self):
        _rt.synchronize()


class _StreamNS:
    null = _Stream()


class cuda:
    Stream = _StreamNS
