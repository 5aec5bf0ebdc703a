"""Seeded random parity cases for the per-node C ABI, runnable on any backend of harness.py.

Each `run_*` builds its inputs in the backend's memory space from a numpy description, calls the
C ABI and returns plain numpy outputs, so two backends can be compared byte for byte.
Inputs avoid what C/C++ leave undefined (integer division by zero, float->int out of range),
because there the reference's HOST build, its DEVICE build and any restatement may legitimately
differ.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import harness as H
from aresdb_b200 import cabi as A

COLUMN_TYPES = [A.Bool, A.Int8, A.Uint8, A.Int16, A.Uint16, A.Int32, A.Uint32, A.Float32]
SCRATCH_TYPES = [A.Int32, A.Uint32, A.Float32]
DIM_TYPES = [A.Bool, A.Int8, A.Uint8, A.Int16, A.Uint16, A.Int32, A.Uint32, A.Float32, A.Int64]
MEASURE_TYPES = [A.Int32, A.Uint32, A.Float32, A.Int64, A.Float64]
UNARY_FNS = [A.Negate, A.Not, A.BitwiseNot, A.IsNull, A.IsNotNull, A.Noop, A.GetWeekStart, A.GetMonthStart,
             A.GetQuarterStart, A.GetYearStart, A.GetDayOfMonth, A.GetDayOfYear, A.GetMonthOfYear,
             A.GetQuarterOfYear, A.GetHLLValue]
BINARY_FNS = [A.And, A.Or, A.Equal, A.NotEqual, A.LessThan, A.LessThanOrEqual, A.GreaterThan,
              A.GreaterThanOrEqual, A.Plus, A.Minus, A.Multiply, A.Divide, A.Mod, A.BitwiseAnd, A.BitwiseOr,
              A.BitwiseXor, A.Floor]
DATE_FNS = set(range(A.GetWeekStart, A.GetQuarterOfYear + 1))
INT_ONLY_BIN = {A.Mod, A.BitwiseAnd, A.BitwiseOr, A.BitwiseXor, A.Floor}


def random_values(rng, data_type, n, nonzero=False, small=False):
    if data_type == A.Bool:
        return np.ones(n, np.uint8) if nonzero else rng.integers(0, 2, n).astype(np.uint8)
    if data_type == A.Float32:
        v = (rng.integers(-400, 400, n) / 8.0).astype(np.float32)
        if small:
            v = np.abs(v)
        if nonzero:
            v[v == 0] = 1.5
        return v
    np_t = H._NP_OF[data_type]
    info = np.iinfo(np_t)
    lo, hi = max(info.min, -100), min(info.max, 100)
    if small:
        lo = max(lo, 0)
    v = rng.integers(lo, hi + 1, n).astype(np_t)
    if nonzero:
        v[v == 0] = 3
        v[v == -1] = 5 if info.min < 0 else v[v == -1]
    return v


class InputSpec:
    """Description of one InputVector, independent of the memory space."""

    def __init__(self, kind, data_type=None, values=None, valid=None, mode=2, start_bit=0, counts=None,
                 const=None, const_valid=True, default=None, default_valid=False):
        self.kind, self.data_type, self.values, self.valid = kind, data_type, values, valid
        self.mode, self.start_bit, self.counts = mode, start_bit, counts
        self.const, self.const_valid = const, const_valid
        self.default, self.default_valid = default, default_valid

    def build(self, be, keep):
        if self.kind == "const":
            return A.const_input(self.const, self.const_valid)
        if self.kind == "scratch":
            buf, noff = H.make_scratch(be, self.data_type, self.values, self.valid)
            keep.append(buf)
            return A.scratch_input(buf.ptr, noff, self.data_type)
        if self.mode == 0:
            dv = A.make_default_value(self.default_valid, self.default, self.data_type)
            return A.vp_input(A.make_vp_slice(None, 0, 0, 0, self.data_type, 0, dv))
        valid = None if self.mode == 1 else self.valid
        counts = self.counts if self.mode == 3 else None
        buf, vp = H.make_column(be, self.data_type, self.values, valid=valid, counts=counts,
                                start_bit=self.start_bit)
        keep.append(buf)
        return A.vp_input(vp)


def random_input(rng, n, allow_const=True, nonzero=False, small=False, force_type=None, rle=None):
    """rle: None, or (run_starts u32[runs+1]) describing a mode-3 column over `rows` rows."""
    kinds = ["column"] * 5 + ["scratch"] * 2 + (["const"] if allow_const else [])
    kind = kinds[rng.integers(0, len(kinds))]
    if kind == "const":
        if rng.integers(0, 2):
            v = float(rng.integers(1 if nonzero else -20, 20)) + 0.5
            if small:
                v = abs(v)
            return InputSpec("const", const=v, const_valid=bool(rng.integers(0, 8) != 0))
        v = int(rng.integers(1 if (nonzero or small) else -20, 20))
        if nonzero and v in (0, -1):
            v = 7
        return InputSpec("const", const=v, const_valid=bool(rng.integers(0, 8) != 0))
    if kind == "scratch":
        dt = force_type if force_type in SCRATCH_TYPES else SCRATCH_TYPES[rng.integers(0, 3)]
        return InputSpec("scratch", dt, random_values(rng, dt, n, nonzero, small), rng.integers(0, 4, n) != 0)
    dt = force_type if force_type is not None else COLUMN_TYPES[rng.integers(0, len(COLUMN_TYPES))]
    mode = int(rng.integers(0, 4)) if rle is not None else int(rng.integers(0, 3))
    if mode == 0:
        d = random_values(rng, dt, 1, nonzero, small)[0]
        return InputSpec("column", dt, mode=0, default=d.item(), default_valid=bool(rng.integers(0, 3) != 0))
    start_bit = int(rng.integers(0, 8))
    if mode == 3:
        runs = len(rle) - 1
        return InputSpec("column", dt, random_values(rng, dt, runs, nonzero, small), rng.integers(0, 4, runs) != 0,
                         mode=3, start_bit=start_bit, counts=rle)
    return InputSpec("column", dt, random_values(rng, dt, n, nonzero, small), rng.integers(0, 4, n) != 0,
                     mode=mode, start_bit=start_bit)


def input_kind_class(spec: InputSpec):
    """'f' float, 's' signed, 'u' unsigned/bool — to steer clear of undefined conversions."""
    if spec.kind == "const":
        return "f" if isinstance(spec.const, float) else "s"
    if spec.data_type == A.Float32:
        return "f"
    return "s" if spec.data_type in (A.Int8, A.Int16, A.Int32) else "u"


def run_transform(be, ins, fn, sink, n, index=None, base_counts=None, start_count=0):
    """sink: ('scratch', dt) | ('dim', dt) | ('measure', dt, agg).  Returns dict of outputs."""
    keep = []
    ivs = [s.build(be, keep) for s in ins]
    idx = be.put(np.arange(n, dtype=np.uint32) if index is None else np.asarray(index, np.uint32))
    bc = be.put(np.asarray(base_counts, np.uint32)) if base_counts is not None else None
    bcp = bc.ptr if bc is not None else None
    width = max(A.DATA_TYPE_BYTES[sink[1]], 1)
    if sink[0] == "scratch":
        noff = H.align(4 * n, 8)
        out = be.zeros(noff + n)
        ov = A.scratch_output(out.ptr, noff, sink[1])
        width = 4
    elif sink[0] == "dim":
        noff = H.align(width * n, 8)
        out = be.zeros(noff + n)
        ov = A.dimension_output(out.ptr, out.at(noff), sink[1])
    else:
        noff = None
        out = be.zeros(width * n)
        ov = A.measure_output(out.ptr, sink[1], sink[2])
    if len(ivs) == 1:
        be.lib.UnaryTransform(ivs[0], ov, idx.ptr, n, bcp, start_count, fn, be.space.stream, be.device)
    else:
        be.lib.BinaryTransform(ivs[0], ivs[1], ov, idx.ptr, n, bcp, start_count, fn, be.space.stream, be.device)
    res = {"values": out.get(np.uint8, width * n)}
    if noff is not None:
        res["valid"] = out.get(np.uint8, n, noff)
    return res


def run_filter(be, ins, fn, n, index=None, base_counts=None, start_count=0):
    keep = []
    ivs = [s.build(be, keep) for s in ins]
    idx = be.put(np.arange(n, dtype=np.uint32) if index is None else np.asarray(index, np.uint32))
    bc = be.put(np.asarray(base_counts, np.uint32)) if base_counts is not None else None
    bcp = bc.ptr if bc is not None else None
    pred = be.zeros(n)
    if len(ivs) == 1:
        m = be.lib.UnaryFilter(ivs[0], idx.ptr, pred.ptr, n, None, 0, bcp, start_count, fn, be.space.stream, be.device)
    else:
        m = be.lib.BinaryFilter(ivs[0], ivs[1], idx.ptr, pred.ptr, n, None, 0, bcp, start_count, fn,
                                be.space.stream, be.device)
    return {"count": m, "index": idx.get(np.uint32, m)}


# ---- dimension blocks ---------------------------------------------------------------------------
def random_dim_block(rng, nd, capacity, n, cardinality=4, null_rate=0.1):
    """Column-major DimensionVector block with few distinct values per dim (so that groups form)."""
    offs, nulls, widths, total = H.dim_layout(nd, capacity)
    block = np.zeros(total, np.uint8)
    for o, no, w in zip(offs, nulls, widths):
        vals = rng.integers(0, cardinality, n).astype(np.uint64)
        valid = (rng.random(n) >= null_rate).astype(np.uint8)
        vals = vals * valid  # store 0 under NULL so that NULL rows with equal validity group together
        col = np.zeros((n, w), np.uint8)
        col[:, :min(w, 8)] = vals.view(np.uint8).reshape(n, 8)[:, :min(w, 8)]
        block[o:o + n * w] = col.reshape(-1)
        block[no:no + n] = valid
    return block


def unpack_dim_rows(block, nd, capacity, n):
    """Rows of a dim block as a list of bytes objects (values in layout order + validity bytes)."""
    offs, nulls, widths, _ = H.dim_layout(nd, capacity)
    rows = []
    for i in range(n):
        parts = [bytes(block[o + i * w:o + (i + 1) * w]) for o, w in zip(offs, widths)]
        parts += [bytes(block[no + i:no + i + 1]) for no in nulls]
        rows.append(b"".join(parts))
    return rows


MEASURE_NP = {A.AGGR_SUM_UNSIGNED: {4: np.uint32, 8: np.uint64}, A.AGGR_SUM_SIGNED: {4: np.int32, 8: np.int64},
              A.AGGR_SUM_FLOAT: {4: np.float32, 8: np.float64}, A.AGGR_MIN_UNSIGNED: {4: np.uint32},
              A.AGGR_MIN_SIGNED: {4: np.int32}, A.AGGR_MIN_FLOAT: {4: np.float32},
              A.AGGR_MAX_UNSIGNED: {4: np.uint32}, A.AGGR_MAX_SIGNED: {4: np.int32},
              A.AGGR_MAX_FLOAT: {4: np.float32}}


def random_measures(rng, agg, value_bytes, n):
    np_t = MEASURE_NP[agg][value_bytes]
    if np.issubdtype(np_t, np.floating):
        return (rng.integers(0, 6400, n) / 64.0).astype(np_t)  # exact in fp32/fp64 in any summation order
    if np.issubdtype(np_t, np.signedinteger):
        return rng.integers(-1000, 1000, n).astype(np_t)
    return rng.integers(0, 2000, n).astype(np_t)


def run_sort_reduce(be, block, nd, capacity, n, measures, value_bytes, agg, index=None):
    db = be.put(block)
    hv = be.zeros(8 * capacity)
    idx = be.put(np.arange(n, dtype=np.uint32) if index is None else np.asarray(index, np.uint32))
    mv = be.put(measures)
    od, oh, oi, om = be.zeros(len(block)), be.zeros(8 * capacity), be.zeros(4 * capacity), be.zeros(value_bytes * capacity)
    kin = A.make_dimension_vector(db.ptr, hv.ptr, idx.ptr, nd, capacity)
    kout = A.make_dimension_vector(od.ptr, oh.ptr, oi.ptr, nd, capacity)
    be.lib.Sort(kin, n, be.space.stream, be.device)
    sorted_hash, sorted_idx = hv.get(np.uint64, n), idx.get(np.uint32, n)
    g = be.lib.Reduce(kin, mv.ptr, kout, om.ptr, value_bytes, n, agg, be.space.stream, be.device)
    return {"g": g, "hash": sorted_hash, "sorted_index": sorted_idx, "out_index": oi.get(np.uint32, g),
            "measures": om.get(np.uint8, g * value_bytes), "dims": od.get(np.uint8),
            "rows": unpack_dim_rows(od.get(np.uint8), nd, capacity, g)}


def run_hash_reduce(be, block, nd, capacity, n, measures, value_bytes, agg):
    db = be.put(block)
    mv = be.put(measures)
    od, om = be.zeros(len(block)), be.zeros(value_bytes * capacity)
    g = be.lib.HashReduce(A.make_dimension_vector(db.ptr, None, None, nd, capacity), mv.ptr,
                          A.make_dimension_vector(od.ptr, None, None, nd, capacity), om.ptr, value_bytes, n, agg,
                          be.space.stream, be.device)
    rows = unpack_dim_rows(od.get(np.uint8), nd, capacity, g)
    meas = om.get(np.uint8, g * value_bytes).reshape(g, value_bytes)
    return {"g": g, "groups": {r: bytes(m) for r, m in zip(rows, meas)}}


DIM_CONFIGS = [(0, 0, 0, 0, 1), (0, 0, 1, 0, 0), (0, 0, 1, 1, 1), (0, 1, 0, 1, 0), (1, 0, 1, 0, 2), (0, 0, 2, 1, 0),
               (0, 1, 2, 2, 2), (1, 1, 0, 0, 0)]
