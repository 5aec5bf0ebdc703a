"""Test harness: memory spaces + loaders for the three implementations of the C ABI.

TEST INFRASTRUCTURE ONLY.  Nothing under aresdb_b200/ imports this module or anything
under oracle/.
"""
from __future__ import annotations

import ctypes as C
import functools
from pathlib import Path

import numpy as np
import pytest

from aresdb_b200 import cabi

ROOT = Path(__file__).resolve().parent.parent
REF_DIR = ROOT / "oracle" / "_ref"
ORACLE_LIB = ROOT / "oracle" / "build" / "liboracle.so"


class Buf:
    """A byte buffer in the memory space of a backend."""

    def __init__(self, space, handle, nbytes):
        self.space, self.handle, self.nbytes = space, handle, nbytes

    @property
    def ptr(self) -> int:
        return self.space.ptr_of(self.handle)

    def at(self, offset: int) -> int:
        return self.ptr + offset

    def get(self, dtype=np.uint8, count: int | None = None, offset: int = 0) -> np.ndarray:
        raw = self.space.download(self.handle)
        item = np.dtype(dtype).itemsize
        if count is None:
            count = (self.nbytes - offset) // item
        return raw[offset:offset + count * item].view(dtype).copy()


class HostSpace:
    """Plain host memory (the reference's HOST build treats host pointers as device pointers)."""
    device = 0
    stream = None

    def zeros(self, nbytes: int) -> Buf:
        arr = np.zeros(max(int(nbytes), 1) + 64, dtype=np.uint8)
        off = (-arr.ctypes.data) % 64
        view = arr[off:off + max(int(nbytes), 1)]
        return Buf(self, (arr, view), int(nbytes))

    def put(self, data) -> Buf:
        data = np.ascontiguousarray(data)
        raw = data.view(np.uint8).reshape(-1)
        b = self.zeros(raw.size)
        b.handle[1][:raw.size] = raw
        return b

    def ptr_of(self, handle) -> int:
        return handle[1].ctypes.data

    def download(self, handle) -> np.ndarray:
        return handle[1]

    def sync(self):
        pass


class CudaSpace:
    """Device memory owned by torch (plumbing only); pointers go through the C ABI."""

    def __init__(self, device: int = 0):
        import torch
        self.torch = torch
        self.device = device
        self.stream = None  # legacy default stream, like the reference's unit tests

    def zeros(self, nbytes: int) -> Buf:
        t = self.torch.zeros(max(int(nbytes), 1), dtype=self.torch.uint8, device=f"cuda:{self.device}")
        return Buf(self, t, int(nbytes))

    def put(self, data) -> Buf:
        data = np.ascontiguousarray(data)
        raw = data.view(np.uint8).reshape(-1)
        t = self.torch.from_numpy(raw.copy()).to(f"cuda:{self.device}")
        if t.numel() == 0:
            t = self.torch.zeros(1, dtype=self.torch.uint8, device=f"cuda:{self.device}")
        return Buf(self, t, raw.size)

    def ptr_of(self, handle) -> int:
        return handle.data_ptr()

    def download(self, handle) -> np.ndarray:
        self.torch.cuda.synchronize(self.device)
        return handle.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize(self.device)


class Backend:
    def __init__(self, name: str, lib: cabi.Library, space):
        self.name, self.lib, self.space = name, lib, space
        self.is_gpu = isinstance(space, CudaSpace)

    # convenience
    @property
    def device(self):
        return self.space.device

    def put(self, data) -> Buf:
        return self.space.put(data)

    def zeros(self, nbytes) -> Buf:
        return self.space.zeros(nbytes)


@functools.lru_cache(maxsize=None)
def get_backend(name: str) -> Backend:
    if name == "ref":
        alg, mem = REF_DIR / "libalgorithm.so", REF_DIR / "libmem.so"
        if not alg.exists():
            pytest.skip("oracle/_ref not built (needs /root/reference; run oracle/build_ref.sh)")
        return Backend("ref", cabi.Library(alg, mem, has_plan_api=False, name="ref"), HostSpace())
    if name == "oracle":
        if not ORACLE_LIB.exists():
            import subprocess, sys
            subprocess.run([sys.executable, str(ROOT / "oracle" / "build_oracle.py")], check=True)
        return Backend("oracle", cabi.Library(ORACLE_LIB, None, has_plan_api=False, name="oracle"), HostSpace())
    if name == "b200":
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        lib = cabi.load_engine()
        torch.zeros(1, device="cuda:0")  # create the context before the first C-ABI call
        return Backend("b200", lib, CudaSpace(0))
    raise KeyError(name)


# ---- column / vector builders shared by the tests ---------------------------------------------
def align(n: int, a: int = 8) -> int:
    return (n + a - 1) // a * a


def pack_bits(bits, start_bit: int = 0) -> np.ndarray:
    bits = np.asarray(bits, dtype=np.uint8)
    padded = np.concatenate([np.zeros(start_bit, np.uint8), bits])
    return np.packbits(padded, bitorder="little")


_NP_OF = {cabi.Int8: np.int8, cabi.Uint8: np.uint8, cabi.Int16: np.int16, cabi.Uint16: np.uint16,
          cabi.Int32: np.int32, cabi.Uint32: np.uint32, cabi.Float32: np.float32, cabi.Int64: np.int64,
          cabi.Uint64: np.uint64}


def make_column(be: Backend, data_type: int, values, valid=None, counts=None, start_bit: int = 0,
                default: cabi.DefaultValue | None = None, value_align: int = 64):
    """Builds [counts][nulls][values] like memstore hands it to the query path and returns
    (Buf, VectorPartySlice).  valid=None -> mode 1; counts given -> mode 3."""
    n = len(values)
    if data_type == cabi.Bool:
        vbytes = pack_bits(np.asarray(values, dtype=np.uint8) != 0, start_bit)
    elif data_type == cabi.UUID:
        vbytes = np.ascontiguousarray(values, dtype=np.uint64).view(np.uint8).reshape(-1)
    else:
        vbytes = np.ascontiguousarray(values, dtype=_NP_OF[data_type]).view(np.uint8).reshape(-1)
    parts, nulls_off, values_off = [], 0, 0
    pos = 0
    if counts is not None:
        cb = np.ascontiguousarray(counts, dtype=np.uint32).view(np.uint8)
        parts.append((pos, cb))
        pos = align(pos + cb.size, value_align)
    if valid is not None or counts is not None:
        v = np.ones(n, np.uint8) if valid is None else np.asarray(valid, dtype=np.uint8)
        nb = pack_bits(v != 0, start_bit)
        nulls_off = pos
        parts.append((pos, nb))
        pos = align(pos + nb.size, value_align)
    values_off = pos
    parts.append((pos, vbytes))
    total = pos + vbytes.size
    raw = np.zeros(total, np.uint8)
    for off, b in parts:
        raw[off:off + b.size] = b
    buf = be.put(raw)
    if counts is None and valid is None:
        vp = cabi.make_vp_slice(buf.ptr, 0, 0, start_bit, data_type, n, default)          # mode 1
    elif counts is None:
        vp = cabi.make_vp_slice(buf.ptr, 0, values_off, start_bit, data_type, n, default)  # mode 2
    else:
        vp = cabi.make_vp_slice(buf.ptr, nulls_off, values_off, start_bit, data_type, n, default)  # mode 3
    return buf, vp


def make_scratch(be: Backend, data_type: int, values, valid):
    """Scratch-space vector: values (4 B each) followed, 8-byte aligned, by bool bytes."""
    vb = np.ascontiguousarray(values, dtype=_NP_OF[data_type]).view(np.uint8).reshape(-1)
    nulls_off = align(vb.size, 8)
    raw = np.zeros(nulls_off + len(valid), np.uint8)
    raw[:vb.size] = vb
    raw[nulls_off:] = np.asarray(valid, dtype=np.uint8)
    buf = be.put(raw)
    return buf, nulls_off


def dim_layout(num_dims_per_width, capacity: int):
    """(value offsets, null offsets, total bytes) of a DimensionVector block
    (reference query/common/dimval.go:122-145)."""
    offs, pos = [], 0
    widths = []
    for w, cnt in zip(cabi.DIM_WIDTHS, num_dims_per_width):
        for _ in range(cnt):
            offs.append(pos)
            widths.append(w)
            pos += w * capacity
    nulls = []
    for _ in widths:
        nulls.append(pos)
        pos += capacity
    return offs, nulls, widths, pos
