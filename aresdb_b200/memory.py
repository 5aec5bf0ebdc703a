"""Memory spaces the host-side executors allocate from.

`CudaSpace` hands out torch-owned device buffers (PyTorch is only the device-memory / stream /
torch.distributed plumbing here); `HostSpace` is plain host memory, which is what a HOST-mode
build of the C ABI (the reference's, or a checker) treats as "device" pointers.
"""
from __future__ import annotations

import numpy as np


class Buf:
    """A byte buffer living in a memory space."""

    def __init__(self, space, handle, nbytes):
        self.space, self.handle, self.nbytes = space, handle, nbytes

    @property
    def ptr(self) -> int:
        return self.space.ptr_of(self.handle)

    def at(self, offset: int) -> int:
        return self.ptr + offset

    def get(self, dtype=np.uint8, count: int | None = None, offset: int = 0) -> np.ndarray:
        item = np.dtype(dtype).itemsize
        if count is None:
            count = (self.nbytes - offset) // item
        raw = self.space.download(self.handle, offset, count * item)
        return raw.view(dtype).copy()


class HostSpace:
    device = 0
    stream = None
    is_cuda = False

    def zeros(self, nbytes: int) -> Buf:
        n = max(int(nbytes), 1)
        arr = np.zeros(n + 64, dtype=np.uint8)
        off = (-arr.ctypes.data) % 64
        return Buf(self, (arr, arr[off:off + n]), int(nbytes))

    def put(self, data) -> Buf:
        raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        b = self.zeros(raw.size)
        b.handle[1][:raw.size] = raw
        return b

    def ptr_of(self, handle) -> int:
        return handle[1].ctypes.data

    def download(self, handle, offset=0, nbytes=None) -> np.ndarray:
        v = handle[1]
        return v[offset:] if nbytes is None else v[offset:offset + nbytes]

    def copy(self, dst: Buf, dst_off: int, src: Buf, src_off: int, nbytes: int):
        if nbytes > 0:
            dst.handle[1][dst_off:dst_off + nbytes] = src.handle[1][src_off:src_off + nbytes]

    def sync(self):
        pass


class CudaSpace:
    is_cuda = True

    def __init__(self, device: int = 0, stream: int | None = None):
        import torch
        self.torch = torch
        self.device = device
        self.stream = stream  # raw cudaStream_t (int) or None for the legacy default stream
        self.dev = torch.device(f"cuda:{device}")

    def zeros(self, nbytes: int) -> Buf:
        t = self.torch.zeros(max(int(nbytes), 1), dtype=self.torch.uint8, device=self.dev)
        return Buf(self, t, int(nbytes))

    def empty(self, nbytes: int) -> Buf:
        t = self.torch.empty(max(int(nbytes), 1), dtype=self.torch.uint8, device=self.dev)
        return Buf(self, t, int(nbytes))

    def put(self, data) -> Buf:
        raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if raw.size == 0:
            return self.zeros(0)
        t = self.torch.from_numpy(raw.copy()).to(self.dev)
        return Buf(self, t, raw.size)

    def ptr_of(self, handle) -> int:
        return handle.data_ptr()

    def download(self, handle, offset=0, nbytes=None) -> np.ndarray:
        self.torch.cuda.synchronize(self.device)
        t = handle[offset:] if nbytes is None else handle[offset:offset + nbytes]
        return t.cpu().numpy()

    def copy(self, dst: Buf, dst_off: int, src: Buf, src_off: int, nbytes: int):
        if nbytes > 0:
            dst.handle[dst_off:dst_off + nbytes].copy_(src.handle[src_off:src_off + nbytes])

    def sync(self):
        self.torch.cuda.synchronize(self.device)
