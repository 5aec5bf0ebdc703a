"""Building VectorParty column slices the way memstore hands them to the query path:
one allocation [counts u32 x (len+1)] [null bitmap] [values], each part 64-byte aligned
(reference memstore/vectors/vector.go:261-272), described by a VectorPartySlice
(makeVectorPartySlice, reference query/time_series_aggregate.go:166-206)."""
from __future__ import annotations

import numpy as np

from . import cabi as A

NP_OF = {A.Int8: np.int8, A.Uint8: np.uint8, A.Int16: np.int16, A.Uint16: np.uint16, A.Int32: np.int32,
         A.Uint32: np.uint32, A.Float32: np.float32, A.Int64: np.int64, A.Uint64: np.uint64}


def align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


def pack_bits(bits, start_bit: int = 0) -> np.ndarray:
    bits = np.asarray(bits, dtype=np.uint8)
    if start_bit:
        bits = np.concatenate([np.zeros(start_bit, np.uint8), bits])
    return np.packbits(bits, bitorder="little")


def column_bytes(data_type: int, values, valid=None, counts=None, start_bit: int = 0, value_align: int = 64):
    """Returns (raw bytes, nulls_offset, values_offset, mode)."""
    n = len(values)
    if data_type == A.Bool:
        vbytes = pack_bits(np.asarray(values, dtype=np.uint8) != 0, start_bit)
    elif data_type == A.UUID:
        vbytes = np.ascontiguousarray(values, dtype=np.uint64).view(np.uint8).reshape(-1)
    else:
        vbytes = np.ascontiguousarray(values, dtype=NP_OF[data_type]).view(np.uint8).reshape(-1)
    parts, nulls_off, pos = [], 0, 0
    if counts is not None:
        cb = np.ascontiguousarray(counts, dtype=np.uint32).view(np.uint8)
        parts.append((pos, cb))
        pos = align(pos + cb.size, value_align)
    if valid is not None or counts is not None:
        v = np.ones(n, np.uint8) if valid is None else np.asarray(valid, dtype=np.uint8)
        nb = pack_bits(v != 0, start_bit)
        nulls_off = pos
        parts.append((pos, nb))
        pos = align(pos + nb.size + 1, value_align)
    values_off = pos
    parts.append((pos, vbytes))
    raw = np.zeros(align(pos + vbytes.size + 1, value_align), np.uint8)
    for off, b in parts:
        raw[off:off + b.size] = b
    mode = 1 if (valid is None and counts is None) else (2 if counts is None else 3)
    return raw, nulls_off, values_off, mode


def make_column(space, data_type: int, values, valid=None, counts=None, start_bit: int = 0,
                default: A.DefaultValue | None = None, value_align: int = 64):
    """Uploads a column into `space`; returns (Buf, VectorPartySlice).
    valid=None & counts=None -> mode 1; valid given -> mode 2; counts given -> mode 3."""
    raw, nulls_off, values_off, mode = column_bytes(data_type, values, valid, counts, start_bit, value_align)
    buf = space.put(raw)
    return buf, slice_of(buf.ptr, data_type, len(values), nulls_off, values_off, mode, start_bit, default)


def slice_of(base_ptr: int, data_type: int, length: int, nulls_off: int, values_off: int, mode: int,
             start_bit: int = 0, default: A.DefaultValue | None = None) -> A.VectorPartySlice:
    if mode == 1:
        return A.make_vp_slice(base_ptr, 0, 0, start_bit, data_type, length, default)
    if mode == 2:
        return A.make_vp_slice(base_ptr, 0, values_off, start_bit, data_type, length, default)
    return A.make_vp_slice(base_ptr, nulls_off, values_off, start_bit, data_type, length, default)


def constant_column(data_type: int, value, valid: bool = True) -> A.VectorPartySlice:
    """Mode-0 column: no vectors, every row is the default value."""
    return A.make_vp_slice(None, 0, 0, 0, data_type, 0, A.make_default_value(valid, value, data_type))
