"""The `application/hll` wire format: what the reference returns for an hll query when the client asks
for the register sets instead of the estimates — the carried outputs of the last HyperLogLog call (dimension block, register
counts, sparse / dense register vectors) behind a self-describing header — and its parser.

Layout (reference query/common/hll.go:44-68; writer :853-942 SerializeHeader, query/hll.go:27-110 SerializeHLL; reader
:364-501 parseTimeseriesHLLResult, :547-581 readHLL; container :944-1000 HLLQueryResults, :583-633 ParseHLLQueryResults):

    container   [u32 magic 0xACED0102][u32 padding] then per query  [u32 size][u8 0 = result | 1 = error][3 bytes padding][payload]
    payload     [u8 enum columns][u8 x 5 dims per width][pad 8] [u32 result size][u32 padded dim-vector bytes]
                [u8 vector index per dimension][pad 8] [u32 data type per dimension][pad 8]
                per enum column: [u32 bytes][u16 dimension][u16 pad] names, each followed by "\\0\\n" [pad 8]
                dimension block of `result size` rows [pad 8]   u16 register counts [pad 8]   register vectors [pad 8]

Data types on the wire are the memstore's codes (memstore/common/data_type.go:44-58: published constants).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from . import cabi as A
from .postprocess import DimensionMeta, NULL_STRING, read_dimension
from .query import HLL_REGISTERS, HLLResult

HLL_DATA_HEADER = 0xACED0102
ENUM_DELIMITER = b"\x00\n"
DENSE_THRESHOLD = HLL_REGISTERS // 4

# memstore data type codes (width in bits in the low half, base type above it)
MEM_BOOL, MEM_INT8, MEM_UINT8, MEM_INT16, MEM_UINT16 = 0x00000001, 0x00010008, 0x00020008, 0x00030010, 0x00040010
MEM_INT32, MEM_UINT32, MEM_FLOAT32, MEM_SMALL_ENUM, MEM_BIG_ENUM = 0x00050020, 0x00060020, 0x00070020, 0x00080008, 0x00090010
MEM_UUID, MEM_INT64 = 0x000A0080, 0x000D0040
_MEM_TO_VALUE_TYPE = {MEM_BOOL: A.Bool, MEM_INT8: A.Int8, MEM_UINT8: A.Uint8, MEM_INT16: A.Int16, MEM_UINT16: A.Uint16,
                      MEM_INT32: A.Int32, MEM_UINT32: A.Uint32, MEM_FLOAT32: A.Float32, MEM_SMALL_ENUM: A.Uint8,
                      MEM_BIG_ENUM: A.Uint16, MEM_UUID: A.UUID, MEM_INT64: A.Int64}
_NUMPY = {A.Bool: np.uint8, A.Int8: np.int8, A.Uint8: np.uint8, A.Int16: np.int16, A.Uint16: np.uint16, A.Int32: np.int32,
          A.Uint32: np.uint32, A.Float32: np.float32, A.Int64: np.int64}


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def enum_cases_bytes(cases: list) -> int:
    """CalculateEnumCasesBytes: names + one delimiter each, rounded up to 8."""
    return _pad8(sum(len(c.encode()) for c in cases) + 2 * len(cases))


def dimension_offsets(num_dims_per_width, vector_index: int, length: int) -> tuple:
    """(value offset, validity offset) of the dimension at position `vector_index` of a block of `length` rows
    (GetDimensionStartOffsets, query/common/dimval.go:122-144)."""
    value, start = 0, 0
    widths = [1 << (len(num_dims_per_width) - 1 - i) for i in range(len(num_dims_per_width))]
    for w, cnt in zip(widths, num_dims_per_width):
        if start + cnt > vector_index:
            value += (vector_index - start) * length * w
            break
        start += cnt
        value += cnt * length * w
    value_bytes = sum(w * cnt for w, cnt in zip(widths, num_dims_per_width))
    return value, (value_bytes + vector_index) * length


def header_and_total_size(num_dims_per_width, dim_indexes, data_types, enum_dicts: dict, padded_dim_bytes: int,
                          result_size: int, padded_hll_bytes: int) -> tuple:
    """HLLData.CalculateSizes."""
    header = _pad8(1 + len(num_dims_per_width)) + 8 + _pad8(len(dim_indexes)) + _pad8(4 * len(data_types))
    header += sum(8 + enum_cases_bytes(c) for c in enum_dicts.values())
    return header, header + padded_dim_bytes + _pad8(2 * result_size) + padded_hll_bytes


def serialize_hll_data(num_dims_per_width, dim_indexes, data_types, enum_dicts: dict, dim_block: bytes, counts, hll_vector: bytes) -> bytes:
    """One query's payload.  `dim_block`: the DimensionVector block laid out for exactly len(counts) rows; `dim_indexes[i]`:
    position of query dimension i in that block; `data_types[i]`: its memstore code; `enum_dicts` {dimension: [names]}."""
    counts = np.asarray(counts, np.uint16)
    n = len(counts)
    padded_dims, padded_hll = _pad8(len(dim_block)), _pad8(len(hll_vector))
    header, total = header_and_total_size(num_dims_per_width, dim_indexes, data_types, enum_dicts, padded_dims, n, padded_hll)
    out = bytearray(total)
    pos = 0

    def put(fmt, *v):
        nonlocal pos
        struct.pack_into(fmt, out, pos, *v)
        pos += struct.calcsize(fmt)

    put("<B", len(enum_dicts))
    out[pos:pos + len(num_dims_per_width)] = bytes(int(c) for c in num_dims_per_width)
    pos = _pad8(pos + len(num_dims_per_width))
    put("<II", n, padded_dims)
    for d in dim_indexes:
        put("<B", d)
    pos = _pad8(pos)
    for t in data_types:
        put("<I", t)
    pos = _pad8(pos)
    for dim, cases in enum_dicts.items():
        nbytes = enum_cases_bytes(cases)
        put("<IHH", nbytes, dim, 0)
        start = pos
        for c in cases:
            raw = c.encode() + ENUM_DELIMITER
            out[pos:pos + len(raw)] = raw
            pos += len(raw)
        pos = start + nbytes
    assert pos == header
    out[pos:pos + len(dim_block)] = dim_block
    pos += padded_dims
    out[pos:pos + 2 * n] = counts.tobytes()
    pos += _pad8(2 * n)
    out[pos:pos + len(hll_vector)] = hll_vector
    return bytes(out)


def serialize_hll_result(result: HLLResult, data_types: list, enum_dicts: dict | None = None, time_dimensions=(), from_offset: int = 0,
                         to_offset: int = 0, dst_switch: int = 0) -> bytes:
    """SerializeHLL for an HLLResult of this engine (AggStateFinalizeHLL / the legacy HyperLogLog call): the dimension
    block as it came back, register counts and vectors.  `time_dimensions`: query dimensions whose numeric value is
    shifted back to an instant (utils.AdjustOffset) when the query ran in a time zone, clamped to [0, 2^32 - 1]."""
    q = result.query
    n = result.groups
    # the dimension block re-laid for exactly `groups` rows (values dimension by dimension in layout order, then validity)
    raw = result.dims._raw_cols
    block = bytearray(b"".join([np.ascontiguousarray(v[:n]).tobytes() for v, _ in raw] +
                               [np.ascontiguousarray(vd[:n]).astype(np.uint8).tobytes() for _, vd in raw]))
    vector_index = [q.dim_order.index(d) for d in range(len(q.dimensions))]
    if time_dimensions and (from_offset or to_offset):
        for d in time_dimensions:
            vo, no = dimension_offsets(q.num_dims_per_width, vector_index[d], n)
            vals = np.frombuffer(block, np.uint32, n, vo).copy()
            ok = np.frombuffer(block, np.uint8, n, no) != 0
            v = vals.astype(np.int64)
            off = np.where((dst_switch > 0) & (v >= dst_switch + to_offset), to_offset, from_offset)
            v = np.clip(v - off, 0, 0xFFFFFFFF)
            vals[ok] = v[ok].astype(np.uint32)
            block[vo:vo + 4 * n] = vals.tobytes()
    return serialize_hll_data(q.num_dims_per_width, vector_index, data_types, enum_dicts or {}, bytes(block), result.counts,
                              np.asarray(result.regs, np.uint8).tobytes())


class HLLQueryResultsWriter:
    """HLLQueryResults: the container of several queries' payloads or errors."""

    def __init__(self):
        self.buf = bytearray(struct.pack("<II", HLL_DATA_HEADER, 0))

    def write_result(self, payload: bytes):
        self.buf += struct.pack("<IB3x", len(payload), 0) + payload

    def write_error(self, message: str):
        raw = message.encode()
        self.buf += struct.pack("<IB3x", len(raw), 1) + raw
        self.buf += bytes((8 - (len(raw) & 7)) & 8)       # the reference's padding expression, as written there

    def get_bytes(self) -> bytes:
        return bytes(self.buf)


@dataclass
class HLL:
    """One register set as the parser returns it (query/common/hll.go HLL): sparse (index, rho) pairs or 16384 dense bytes."""
    non_zero_registers: int
    sparse: list | None = None
    dense: bytes | None = None

    def dense_registers(self) -> np.ndarray:
        if self.dense is not None:
            return np.frombuffer(self.dense, np.uint8).copy()
        out = np.zeros(HLL_REGISTERS, np.uint8)
        for index, rho in self.sparse or []:
            out[index] = rho
        return out

    # ---- what the broker does with register sets of several nodes (query/common/hll.go:148-215, 669-733) -------------
    def to_dense(self):
        """ConvertToDense."""
        if self.dense:
            return
        self.dense = self.dense_registers().tobytes()
        self.sparse = None

    def to_sparse(self) -> bool:
        """ConvertToSparse: only when it is the cheaper form (fewer than a quarter of the registers hit)."""
        if self.non_zero_registers * 4 >= HLL_REGISTERS:
            return False
        if self.sparse is not None:
            return True
        d = np.frombuffer(self.dense, np.uint8)
        self.sparse = [(int(i), int(d[i])) for i in np.nonzero(d)[0]]
        self.dense = None
        return True

    def set(self, index: int, rho: int):
        """Set: one more register (each at most once); switches to dense at a quarter of the registers."""
        self.non_zero_registers += 1
        if self.dense:
            b = bytearray(self.dense)
            b[index] = rho
            self.dense = bytes(b)
            return
        self.sparse = (self.sparse or []) + [(index, rho)]
        if self.non_zero_registers * 4 >= HLL_REGISTERS:
            self.to_dense()

    def merge(self, other: "HLL"):
        """Merge: register-wise maximum; the receiver becomes dense."""
        self.to_dense()
        mine = np.frombuffer(self.dense, np.uint8).copy()
        theirs = other.dense_registers()
        self.non_zero_registers += int(np.count_nonzero((mine == 0) & (theirs != 0)))
        self.dense = np.maximum(mine, theirs).tobytes()

    def encode(self) -> bytes:
        """Encode (cache form): the dense bytes, or 3 bytes per sparse register (index low, index high, rho)."""
        if self.dense:
            return self.dense
        return b"".join(bytes((i & 0xFF, i >> 8, r)) for i, r in self.sparse or [])

    def encode_binary(self) -> bytes:
        """EncodeBinary (wire form): the dense bytes, or 4 bytes per sparse register — `uint32(int8(rho)) << 16 | index`,
        the sign extension of a rho >= 128 included, as the reference writes it."""
        if self.dense:
            return self.dense
        return b"".join(struct.pack("<I", ((((r - 256 if r >= 128 else r) & 0xFFFFFFFF) << 16) & 0xFFFFFFFF) | i) for i, r in self.sparse or [])

    @classmethod
    def decode(cls, data: bytes) -> "HLL":
        """Decode: 16384 bytes are dense registers, anything else 3-byte sparse registers."""
        if len(data) == HLL_REGISTERS:
            return cls(int(np.count_nonzero(np.frombuffer(data, np.uint8))), dense=bytes(data))
        n = len(data) // 3
        return cls(n, sparse=[(data[3 * i] | (data[3 * i + 1] << 8), data[3 * i + 2]) for i in range(n)])

    def compute(self) -> float:
        """Compute: the HyperLogLog++ estimate of the register set."""
        from .postprocess import hll_estimate
        return hll_estimate(self.dense_registers())


def read_hll(vector: bytes, count: int, offset: int) -> tuple:
    """readHLL: `count` below the dense threshold = that many 4-byte (index u16, rho u8) entries, else 16384 bytes."""
    if count < DENSE_THRESHOLD:
        entries = np.frombuffer(vector, "<u4", count, offset)
        return HLL(count, sparse=[(int(e & 0xFFFF), int((e >> 16) & 0xFF)) for e in entries]), offset + 4 * count
    dense = bytes(vector[offset:offset + HLL_REGISTERS])
    return HLL(int(np.count_nonzero(np.frombuffer(dense, np.uint8))), dense=dense), offset + HLL_REGISTERS


def parse_hll_data(buffer: bytes, ignore_enum: bool = False) -> dict:
    """parseTimeseriesHLLResult: payload -> nested {dimension string: ... HLL}."""
    if len(buffer) == 0:
        return {}
    pos = 0
    num_enum = buffer[0]
    per_width = list(buffer[1:6])
    total_dims = sum(per_width)
    pos = _pad8(6)
    result_size, padded_dims = struct.unpack_from("<II", buffer, pos)
    pos += 8
    dim_indexes = list(buffer[pos:pos + total_dims])
    pos += _pad8(total_dims)
    data_types = list(struct.unpack_from(f"<{total_dims}I", buffer, pos))
    for t in data_types:
        if t not in _MEM_TO_VALUE_TYPE:
            raise ValueError(f"invalid data type 0x{t:08x}")
    pos += _pad8(4 * total_dims)
    enum_dicts = {}
    for _ in range(num_enum):
        nbytes, dim = struct.unpack_from("<IH", buffer, pos)
        pos += 8
        cases = bytes(buffer[pos:pos + nbytes]).split(ENUM_DELIMITER)[:-1]   # the last piece is the padding
        enum_dicts[dim] = [c.decode() for c in cases]
        pos += nbytes
    header = pos
    counts = np.frombuffer(buffer, "<u2", result_size, header + padded_dims)
    vector = memoryview(buffer)[header + padded_dims + _pad8(2 * result_size):]
    columns = []
    for d in range(total_dims):
        vo, no = dimension_offsets(per_width, dim_indexes[d], result_size)
        vt = _MEM_TO_VALUE_TYPE[data_types[d]]
        group = next(i for i, c in enumerate(np.cumsum(per_width)) if dim_indexes[d] < c)
        width = 1 << (len(per_width) - 1 - group)                      # bytes per value in the block
        raw = np.frombuffer(buffer, np.uint8, width * result_size, header + vo).reshape(result_size, width)
        if vt == A.UUID:
            vals = raw
        else:
            vals = raw[:, :np.dtype(_NUMPY[vt]).itemsize].copy().view(_NUMPY[vt]).reshape(-1)
        valid = np.frombuffer(buffer, np.uint8, result_size, header + no)
        names = None if ignore_enum else enum_dicts.get(d)
        columns.append((vals, valid, vt, DimensionMeta(enum_names=names) if names is not None else None))
    out: dict = {}
    offset = 0
    for i in range(result_size):
        hll, offset = read_hll(vector, int(counts[i]), offset)
        cur = out
        for d, (vals, valid, vt, meta) in enumerate(columns):
            key = read_dimension(vals[i], bool(valid[i]), vt, meta)
            key = NULL_STRING if key is None else key
            if d == total_dims - 1:
                cur[key] = hll
            else:
                cur = cur.setdefault(key, {})
    return out


def parse_hll_query_results(data: bytes, ignore_enum: bool = False) -> tuple:
    """ParseHLLQueryResults: container -> ([result | None], [error message | None]), one pair per query."""
    magic, = struct.unpack_from("<I", data, 0)
    if magic != HLL_DATA_HEADER:
        raise ValueError(f"header {magic:x} does not match HLLDataHeader {HLL_DATA_HEADER:x}")
    pos, results, errors = 8, [], []
    while pos + 4 <= len(data):
        size, is_err = struct.unpack_from("<IB", data, pos)
        pos += 8
        if pos + size > len(data):
            break
        payload = bytes(data[pos:pos + size])
        pos += size
        if is_err:
            results.append(None)
            errors.append(payload.decode())
        else:
            results.append(parse_hll_data(payload, ignore_enum))
            errors.append(None)
    return results, errors


_MEM_BYTES = {MEM_BOOL: 1, MEM_INT8: 1, MEM_UINT8: 1, MEM_SMALL_ENUM: 1, MEM_INT16: 2, MEM_UINT16: 2, MEM_BIG_ENUM: 2,
              MEM_INT32: 4, MEM_UINT32: 4, MEM_FLOAT32: 4, MEM_INT64: 8}


def _value_bytes(text: str, mem_type: int) -> bytes:
    """ValueFromString for the dimension types a group-by produces."""
    n = _MEM_BYTES[mem_type]
    if mem_type == MEM_FLOAT32:
        return struct.pack("<f", float(text))
    if mem_type == MEM_BOOL:
        return bytes([1 if text.lower() == "true" or text == "1" else 0])
    signed = mem_type in (MEM_INT8, MEM_INT16, MEM_INT32, MEM_INT64)
    return int(text).to_bytes(n, "little", signed=signed)


def build_vectors_from_hll_result(result: dict, data_types: list, enum_dicts: dict, dimension_vector_index: list) -> tuple:
    """BuildVectorsFromHLLResult (query/common/hll.go:1002-1125): a nested {dimension string: ... HLL} result — e.g. what a
    broker holds after merging its nodes' answers — back to (register vector, dimension block, count vector).  Keys are
    visited in string order, children first; enum names go back through `enum_dicts` {dimension: {name: id}}; a register set
    below the dense threshold is written sparse (4 bytes per register), else dense; layout position p of the block holds
    query dimension dimension_vector_index[p]."""
    nd = len(data_types)
    dim_vectors, validity = [bytearray() for _ in range(nd)], [bytearray() for _ in range(nd)]
    hll_vector, count_vector = bytearray(), bytearray()

    def walk(d: int, node) -> int:
        if isinstance(node, HLL):
            count = node.non_zero_registers
            leaf = HLL(node.non_zero_registers, None if node.sparse is None else list(node.sparse), node.dense)
            if count < DENSE_THRESHOLD:
                if not leaf.to_sparse():
                    raise ValueError("Failed to convert HLL to sparse")
            else:
                leaf.to_dense()
            hll_vector.extend(leaf.encode_binary())
            count_vector.extend(struct.pack("<H", count))
            return 1
        if not isinstance(node, dict):
            raise ValueError(f"unknown type {type(node).__name__}")
        width = _MEM_BYTES[data_types[d]]
        size = 0
        for key in sorted(node):
            child = walk(d + 1, node[key])
            if key == NULL_STRING:
                value, ok = bytes(width), 0
            elif d in enum_dicts:
                if width not in (1, 2):
                    raise ValueError(f"data width {width} doesn't match any enum")
                value, ok = int(enum_dicts[d].get(key, 0)).to_bytes(width, "little"), 1
            else:
                value, ok = _value_bytes(key, data_types[d]), 1
            dim_vectors[d].extend(value * child)
            validity[d].extend(bytes([ok]) * child)
            size += child
        return size

    walk(0, result)
    block = b"".join(bytes(dim_vectors[i]) for i in dimension_vector_index) + b"".join(bytes(validity[i]) for i in dimension_vector_index)
    return bytes(hll_vector), block, bytes(count_vector)
