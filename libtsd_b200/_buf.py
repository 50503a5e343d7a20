"""Buffer plumbing between numpy / torch arrays and the C ABI (pointers, strides, memory kind)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DEVICE, HOST, TsdGpuError, bind_torch_stream

try:  # torch is only needed for device-resident buffers
    import torch
except Exception:  # pragma: no cover
    torch = None


def is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class Batch:
    """A [nchan][n] view of a sample buffer: pointer, channel stride (in samples), memory kind."""

    def __init__(self, x, dtype, nchan: int, what: str = "x"):
        self.torch = is_torch(x)
        self.ndim = x.ndim
        if self.torch:
            tdt = {np.complex64: torch.complex64, np.float32: torch.float32}[dtype]
            if not x.is_cuda:
                raise TsdGpuError(f"{what}: torch tensors must live on the GPU (use numpy arrays for host data)")
            if x.dtype != tdt:
                raise TsdGpuError(f"{what}: expected dtype {tdt}, got {x.dtype}")
            if x.ndim == 1:
                x = x.unsqueeze(0)
            if x.ndim != 2 or (x.shape[1] > 1 and x.stride(1) != 1):
                raise TsdGpuError(f"{what}: expected [nchan, n] with contiguous samples")
            self.arr = x
            self.n = int(x.shape[1])
            self.stride = int(x.stride(0)) if x.shape[0] > 1 else max(self.n, 1)
            self.ptr = C.c_void_p(x.data_ptr())
            self.mem = DEVICE
            bind_torch_stream()   # stream-ordered with the producer / consumer kernels of the tensor
        else:
            a = np.asarray(x)
            if a.dtype != dtype:
                a = a.astype(dtype)
            if a.ndim == 1:
                a = a[None, :]
            if a.ndim != 2:
                raise TsdGpuError(f"{what}: expected [nchan, n]")
            if a.shape[1] > 1 and a.strides[1] != a.itemsize:
                a = np.ascontiguousarray(a)
            if a.shape[0] > 1 and (a.strides[0] % a.itemsize != 0 or a.strides[0] < a.shape[1] * a.itemsize):
                a = np.ascontiguousarray(a)
            self.arr = a
            self.n = int(a.shape[1])
            self.stride = int(a.strides[0] // a.itemsize) if a.shape[0] > 1 else max(self.n, 1)
            self.ptr = C.c_void_p(a.ctypes.data)
            self.mem = HOST
        if self.arr.shape[0] != nchan:
            raise TsdGpuError(f"{what}: {self.arr.shape[0]} channels given, the filter was created for {nchan}")


def empty_like_batch(b: Batch, dtype, n: int):
    """Output buffer [nchan, n] of the same kind (numpy / torch-cuda) as the input batch."""
    nchan = b.arr.shape[0]
    if b.torch:
        tdt = {np.complex64: torch.complex64, np.float32: torch.float32}[dtype]
        return torch.empty((nchan, n), dtype=tdt, device=b.arr.device)
    return np.empty((nchan, n), dtype=dtype)


def restore_shape(y, ndim: int):
    return y[0] if ndim == 1 else y
