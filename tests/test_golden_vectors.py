"""Known-answer vectors of the reference's own native unit tests, re-expressed against the C ABI.

Every test names the gtest it restates (reference query/algorithm_unittest.cu:<line>); the input
and expected arrays are the literal values of that test.  They run against all three
implementations of the ABI: the reference HOST build (pins the bindings), the C restatement
(pins the oracle) and the B200 engine (`-m gpu`).
"""
import calendar

import numpy as np
import pytest

import harness as H
from aresdb_b200 import cabi as A

F32 = np.float32


def _scratch_out(be, n, dt=A.Int32):
    buf = be.zeros(16 + n + 8)
    return buf, A.scratch_output(buf.ptr, 16, dt)


def test_abi_struct_sizes():
    """SURVEY.md §8c layout facts (verified there against the reference build with ctypes)."""
    import ctypes
    for cls, size in A.EXPECTED_SIZES.items():
        assert ctypes.sizeof(cls) == size, cls.__name__


def test_unary_transform_check_int(backend):
    """UnaryTransformTest.CheckInt :79 — Negate on an Int32 mode-2 column into scratch."""
    be = backend
    idx = be.put(np.array([0, 1, 2, 0, 0, 0], np.uint32))
    col, vp = H.make_column(be, A.Int32, [-1, 1, 0], valid=[1, 1, 0], value_align=8)
    out, ov = _scratch_out(be, 3)
    be.lib.UnaryTransform(A.vp_input(vp), ov, idx.ptr, 3, None, 0, A.Negate, be.space.stream, be.device)
    assert out.get(np.int32, 3).tolist() == [1, -1, 0]
    assert out.get(np.uint8, 3, 16).tolist() == [1, 1, 0]


def test_unary_transform_check_constant(backend):
    """UnaryTransformTest.CheckConstant :127."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    out, ov = _scratch_out(be, 3)
    be.lib.UnaryTransform(A.const_input(1), ov, idx.ptr, 3, None, 0, A.Negate, be.space.stream, be.device)
    assert out.get(np.int32, 3).tolist() == [-1, -1, -1]
    assert out.get(np.uint8, 3, 16).tolist() == [1, 1, 1]


def test_unary_transform_measure_output_and_avg(backend):
    """UnaryTransformTest.CheckMeasureOutputIteratorForAvg :167 — xcount, identity, avg packing."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    counts = be.put(np.array([0, 3, 9, 12], np.uint32))
    col, vp = H.make_column(be, A.Int32, [-1, 1, 0], valid=[1, 1, 0], value_align=8)
    out = be.zeros(24)
    mo = A.measure_output(out.ptr, A.Int64, A.AGGR_SUM_SIGNED)
    be.lib.UnaryTransform(A.vp_input(vp), mo, idx.ptr, 3, counts.ptr, 0, A.Negate, be.space.stream, be.device)
    assert out.get(np.int64, 3).tolist() == [3, -6, 0]
    be.lib.UnaryTransform(A.vp_input(vp), mo, idx.ptr, 3, counts.ptr, 0, A.Noop, be.space.stream, be.device)
    assert out.get(np.int64, 3).tolist() == [-3, 6, 0]
    mo2 = A.measure_output(out.ptr, A.Float64, A.AGGR_AVG_FLOAT)
    be.lib.UnaryTransform(A.vp_input(vp), mo2, idx.ptr, 3, counts.ptr, 0, A.Noop, be.space.stream, be.device)
    raw = out.get(np.uint32, 6)
    assert raw[0:1].view(F32)[0] == F32(-1.0) and raw[1] == 3
    assert raw[2:3].view(F32)[0] == F32(1.0) and raw[3] == 6
    assert out.get(np.int64, 3)[2] == 0


def test_unary_transform_dimension_output(backend):
    """UnaryTransformTest.CheckDimensionOutputIterator :228 — Int16 dim values + validity bytes."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    col, vp = H.make_column(be, A.Int16, [-1, 1, 0], valid=[1, 1, 0], value_align=8)
    out = be.zeros(9)
    do = A.dimension_output(out.ptr, out.at(6), A.Int16)
    be.lib.UnaryTransform(A.vp_input(vp), do, idx.ptr, 3, None, 0, A.Negate, be.space.stream, be.device)
    assert out.get(np.uint8, 9).tolist() == [1, 0, 0xFF, 0xFF, 0, 0, 1, 1, 0]
    be.lib.UnaryTransform(A.vp_input(vp), do, idx.ptr, 3, None, 0, A.Noop, be.space.stream, be.device)
    assert out.get(np.uint8, 9).tolist() == [0xFF, 0xFF, 1, 0, 0, 0, 1, 1, 0]


def test_unary_filter_with_record_ids(backend):
    """UnaryFilterTest.CheckFilter :271 — IsNotNull on scratch input, RecordID vector zipped."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    pred = be.zeros(3)
    rec = be.put(np.array([0, 0, 0, 1, 0, 2], np.uint32))  # {batchID 0, index i}
    sbuf, noff = H.make_scratch(be, A.Int32, [1, 0, 1], [1, 0, 1])
    import ctypes as C
    recs = (C.c_void_p * 1)(rec.ptr)
    n = be.lib.UnaryFilter(A.scratch_input(sbuf.ptr, noff, A.Int32), idx.ptr, pred.ptr, 3, C.addressof(recs) if not be.is_gpu else _dev_ptr_array(be, [rec.ptr]), 1,
                           None, 0, A.IsNotNull, be.space.stream, be.device)
    assert n == 2
    assert idx.get(np.uint32, 2).tolist() == [0, 2]
    assert rec.get(np.uint32, 4).tolist() == [0, 0, 0, 2]


def _dev_ptr_array(be, ptrs):
    # The reference passes a HOST array of device pointers (Go slice, query/time_series_aggregate.go:378).
    import ctypes as C
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    be._keep = arr
    return C.addressof(arr)


def test_unary_filter_all_empty(backend):
    """UnaryFilterTest.AllEmpty :328 — Negate of zeros keeps nothing."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    pred = be.zeros(3)
    sbuf, noff = H.make_scratch(be, A.Int32, [0, 0, 0], [1, 1, 1])
    n = be.lib.UnaryFilter(A.scratch_input(sbuf.ptr, noff, A.Int32), idx.ptr, pred.ptr, 3, None, 0, None, 0,
                           A.Negate, be.space.stream, be.device)
    assert n == 0


def test_binary_transform_check_int(backend):
    """BinaryTransformTest.CheckInt :356 — Plus with NULLs on either side."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    l, lvp = H.make_column(be, A.Int32, [-1, 1, 0], valid=[1, 1, 0], value_align=8)
    r, rvp = H.make_column(be, A.Int32, [0, 1, -1], valid=[0, 1, 1], value_align=8)
    out, ov = _scratch_out(be, 3)
    be.lib.BinaryTransform(A.vp_input(lvp), A.vp_input(rvp), ov, idx.ptr, 3, None, 0, A.Plus, be.space.stream, be.device)
    assert out.get(np.int32, 3).tolist() == [0, 2, 0]
    assert out.get(np.uint8, 3, 16).tolist() == [0, 1, 0]


def test_binary_transform_float_and_unpacked_bool(backend):
    """BinaryTransformTest.CheckFloatAndUnpackedBoolIter :488 — int32 (-) float32 promotes to float."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    l, lo = H.make_scratch(be, A.Int32, [-1, 1, 0], [1, 1, 1])
    r, ro = H.make_scratch(be, A.Float32, [1.1, -1.1, 0.1], [1, 1, 1])
    out, ov = _scratch_out(be, 3, A.Float32)
    li, ri = A.scratch_input(l.ptr, lo, A.Int32), A.scratch_input(r.ptr, ro, A.Float32)
    be.lib.BinaryTransform(li, ri, ov, idx.ptr, 3, None, 0, A.Minus, be.space.stream, be.device)
    exp = (np.array([-1, 1, 0], F32) - np.array([1.1, -1.1, 0.1], F32))
    assert out.get(F32, 3).tobytes() == exp.tobytes()
    assert out.get(np.uint8, 3, 16).tolist() == [1, 1, 1]
    be.lib.BinaryTransform(li, ri, ov, idx.ptr, 3, None, 0, A.Multiply, be.space.stream, be.device)
    exp = (np.array([-1, 1, 0], F32) * np.array([1.1, -1.1, 0.1], F32))
    assert out.get(F32, 3).tobytes() == exp.tobytes()


def test_binary_transform_constant(backend):
    """BinaryTransformTest.CheckConstantIterator :562 — scratch int32 + ConstFloat."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    l, lo = H.make_scratch(be, A.Int32, [-1, 1, 0], [1, 1, 1])
    out, ov = _scratch_out(be, 3, A.Float32)
    be.lib.BinaryTransform(A.scratch_input(l.ptr, lo, A.Int32), A.const_input(0.1), ov, idx.ptr, 3, None, 0,
                           A.Plus, be.space.stream, be.device)
    exp = np.array([-1, 1, 0], F32) + F32(0.1)
    assert out.get(F32, 3).tobytes() == exp.tobytes()
    assert out.get(np.uint8, 3, 16).tolist() == [1, 1, 1]


def test_binary_filter(backend):
    """BinaryFilterTest.CheckFilter :614 — int32 > float32."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    pred = be.zeros(3)
    l, lo = H.make_scratch(be, A.Int32, [0, 1, 2], [1, 1, 1])
    r, ro = H.make_scratch(be, A.Float32, [0.1, 0.9, 1.9], [1, 1, 1])
    n = be.lib.BinaryFilter(A.scratch_input(l.ptr, lo, A.Int32), A.scratch_input(r.ptr, ro, A.Float32), idx.ptr,
                            pred.ptr, 3, None, 0, None, 0, A.GreaterThan, be.space.stream, be.device)
    assert n == 2
    assert idx.get(np.uint32, 2).tolist() == [1, 2]


def test_binary_transform_measure_output(backend):
    """BinaryTransformTest.CheckMeasureOutputIterator :664 — float measure x RLE count, NULL -> 0."""
    be = backend
    idx = be.put(np.array([0, 1, 2], np.uint32))
    counts = be.put(np.array([0, 3, 9, 10], np.uint32))
    l, lo = H.make_scratch(be, A.Int32, [-1, 1, 0], [1, 1, 0])
    r, ro = H.make_scratch(be, A.Float32, [1.1, -1.1, 0.1], [1, 1, 1])
    out = be.zeros(12)
    mo = A.measure_output(out.ptr, A.Float32, A.AGGR_SUM_FLOAT)
    be.lib.BinaryTransform(A.scratch_input(l.ptr, lo, A.Int32), A.scratch_input(r.ptr, ro, A.Float32), mo, idx.ptr,
                           3, counts.ptr, 0, A.Minus, be.space.stream, be.device)
    d = np.array([-1, 1, 0], F32) - np.array([1.1, -1.1, 0.1], F32)
    exp = np.array([d[0] * F32(3), d[1] * F32(6), 0], F32)
    assert out.get(F32, 3).tobytes() == exp.tobytes()


def test_init_index_vector(backend):
    """InitIndexVectorTest :719."""
    be = backend
    idx = be.zeros(12)
    be.lib.InitIndexVector(idx.ptr, 0, 3, be.space.stream, be.device)
    assert idx.get(np.uint32, 3).tolist() == [0, 1, 2]
    be.lib.InitIndexVector(idx.ptr, 7, 3, be.space.stream, be.device)
    assert idx.get(np.uint32, 3).tolist() == [7, 8, 9]


def test_sort_dim_column_vector(backend):
    """SortDimColumnVectorTest.CheckSort :979 — rows 0 and 2 are equal and must end adjacent."""
    be = backend
    keys = np.zeros(30, np.uint8)
    keys[0:12].view(np.uint32)[:] = [1, 2, 1]
    keys[12:18].view(np.uint16)[:] = [1, 2, 1]
    keys[18:21] = [1, 2, 1]
    kb = be.put(keys)
    idx = be.put(np.array([0, 1, 2], np.uint32))
    hv = be.zeros(24)
    dv = A.make_dimension_vector(kb.ptr, hv.ptr, idx.ptr, (0, 0, 1, 1, 1), 3)
    be.lib.Sort(dv, 3, be.space.stream, be.device)
    assert idx.get(np.uint32, 3).tolist() in ([1, 0, 2], [0, 2, 1])
    h = hv.get(np.uint64, 3)
    assert h[0] <= h[1] <= h[2]


REDUCE_DIMS = [1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 1, 0, 0, 0, 1, 0, 2, 0, 3, 0,
               2, 0, 3, 0, 1, 0, 1, 2, 3, 2, 3, 1] + [1] * 18


def test_reduce_dim_column_vector(backend):
    """ReduceDimColumnVectorTest.CheckReduce :1018 — sum u32 over runs + dim gather layout."""
    be = backend
    ind = be.put(np.array(REDUCE_DIMS, np.uint8))
    ih = be.put(np.array([1, 1, 2, 2, 3, 3], np.uint64))
    ii = be.put(np.array([1, 3, 2, 4, 0, 5], np.uint32))
    iv = be.put(np.array([5, 1, 3, 2, 4, 6], np.uint32))
    od, oh, oi, ov = be.zeros(60), be.zeros(48), be.zeros(24), be.zeros(24)
    nd = (0, 0, 1, 1, 1)
    n = be.lib.Reduce(A.make_dimension_vector(ind.ptr, ih.ptr, ii.ptr, nd, 6), iv.ptr,
                      A.make_dimension_vector(od.ptr, oh.ptr, oi.ptr, nd, 6), ov.ptr, 4, 6, A.AGGR_SUM_UNSIGNED,
                      be.space.stream, be.device)
    assert n == 3
    assert ov.get(np.uint32, 3).tolist() == [3, 7, 11]
    assert oi.get(np.uint32, 3).tolist() == [1, 2, 0]
    exp = [2, 0, 0, 0, 3, 0, 0, 0, 1, 0, 0, 0] + [0] * 12 + [2, 0, 3, 0, 1, 0] + [0] * 6 + [2, 3, 1, 0, 0, 0] + \
          [1, 1, 1, 0, 0, 0] * 3
    assert od.get(np.uint8, 60).tolist() == exp


def test_reduce_by_avg(backend):
    """SortAndReduceTest.CheckReduceByAvg :1086 — packed (avg f32, count u32) rolling average."""
    be = backend
    dims = [1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 1, 0, 0, 0] + [1] * 6
    ind = be.put(np.array(dims, np.uint8))
    ih = be.put(np.array([1, 1, 2, 2, 3, 3], np.uint64))
    ii = be.put(np.array([1, 3, 2, 4, 0, 5], np.uint32))
    vals = np.zeros(12, np.uint32)
    vals[0::2] = np.array([5.0, 1.0, 3.0, 2.0, 4.0, 6.0], F32).view(np.uint32)
    vals[1::2] = 1
    iv = be.put(vals)
    od, oh, oi, ov = be.zeros(30), be.zeros(48), be.zeros(24), be.zeros(48)
    nd = (0, 0, 1, 0, 0)
    n = be.lib.Reduce(A.make_dimension_vector(ind.ptr, ih.ptr, ii.ptr, nd, 6), iv.ptr,
                      A.make_dimension_vector(od.ptr, oh.ptr, oi.ptr, nd, 6), ov.ptr, 8, 6, A.AGGR_AVG_FLOAT,
                      be.space.stream, be.device)
    assert n == 3
    raw = ov.get(np.uint32, 6)
    assert raw[0::2].view(F32).tolist() == [1.5, 3.5, 5.5]
    assert raw[1::2].tolist() == [2, 2, 2]
    assert oi.get(np.uint32, 3).tolist() == [1, 2, 0]
    assert od.get(np.uint8, 30).tolist() == [2, 0, 0, 0, 3, 0, 0, 0, 1, 0, 0, 0] + [0] * 12 + [1, 1, 1, 0, 0, 0]


def test_sort_and_reduce_check_hash(backend):
    """SortAndReduceTest.CheckHash :1160 — pins murmur3 of the packed row AND the output order."""
    be = backend
    dims = be.put(np.array([2, 1, 0, 3, 0, 1, 2, 3] + [1] * 8, np.uint8))
    meas = be.put(np.ones(8, np.uint32))
    hv, idx = be.zeros(64), be.zeros(32)
    od, om, oh, oi = be.zeros(16), be.zeros(32), be.zeros(64), be.zeros(32)
    nd = (0, 0, 0, 0, 1)
    be.lib.InitIndexVector(idx.ptr, 0, 8, be.space.stream, be.device)
    kin = A.make_dimension_vector(dims.ptr, hv.ptr, idx.ptr, nd, 8)
    kout = A.make_dimension_vector(od.ptr, oh.ptr, oi.ptr, nd, 8)
    be.lib.Sort(kin, 8, be.space.stream, be.device)
    n = be.lib.Reduce(kin, meas.ptr, kout, om.ptr, 4, 8, A.AGGR_SUM_UNSIGNED, be.space.stream, be.device)
    assert n == 4
    assert od.get(np.uint8, 16).tolist() == [2, 0, 3, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0]
    assert om.get(np.uint32, 8).tolist() == [2, 2, 2, 2, 0, 0, 0, 0]
    assert oi.get(np.uint32, 8).tolist() == [0, 2, 3, 1, 0, 0, 0, 0]
    # hash known-answers derived through the reference build
    assert [hex(x) for x in hv.get(np.uint64, 8)[::2]] == ['0x60e187b4814392c4', '0x7cb3f5c58dab264c',
                                                           '0xb73e42bb654cee53', '0xca410abc0a9d4c6b']


def _ts(y, m, d):
    return calendar.timegm((y, m, d, 0, 0, 0))


def test_date_functors(backend):
    """DateFunctorsTest.CheckGetStarts :1420 — month / quarter / year start through the C ABI."""
    be = backend
    be.lib.BootstrapDevice()
    idx = be.put(np.array([0, 1, 2], np.uint32))
    sbuf, noff = H.make_scratch(be, A.Uint32, [0, _ts(2018, 6, 11), _ts(1970, 1, 1)], [0, 1, 1])
    out, ov = _scratch_out(be, 3)
    si = A.scratch_input(sbuf.ptr, noff, A.Int32)
    for fn, exp in ((A.GetMonthStart, _ts(2018, 6, 1)), (A.GetQuarterStart, _ts(2018, 4, 1)),
                    (A.GetYearStart, _ts(2018, 1, 1))):
        be.lib.UnaryTransform(si, ov, idx.ptr, 3, None, 0, fn, be.space.stream, be.device)
        assert out.get(np.uint32, 3).tolist() == [0, exp, _ts(1970, 1, 1)]
        assert out.get(np.uint8, 3, 16).tolist() == [0, 1, 1]


def test_hash_reduce(backend):
    """HashReductionTest.CheckReduce :1957 — order-free map compare."""
    be = backend
    ind = be.put(np.array(REDUCE_DIMS, np.uint8))
    iv = be.put(np.array([5, 1, 3, 2, 4, 6], np.uint32))
    od, ov = be.zeros(60), be.zeros(24)
    nd = (0, 0, 1, 1, 1)
    n = be.lib.HashReduce(A.make_dimension_vector(ind.ptr, None, None, nd, 6), iv.ptr,
                          A.make_dimension_vector(od.ptr, None, None, nd, 6), ov.ptr, 4, 6, A.AGGR_SUM_UNSIGNED,
                          be.space.stream, be.device)
    assert n == 3
    d = od.get(np.uint8, 60)
    got = {}
    for i in range(3):
        key = (int(d[0:24].view(np.uint32)[i]), int(d[24:36].view(np.uint16)[i]), int(d[36 + i]),
               int(d[42 + i]), int(d[48 + i]), int(d[54 + i]))
        got[key] = int(ov.get(np.uint32, 3)[i])
    assert got == {(2, 2, 2, 1, 1, 1): 3, (1, 1, 1, 1, 1, 1): 11, (3, 3, 3, 1, 1, 1): 7}


def _free_outputs(be, *ptrs):
    import ctypes as C
    for p in ptrs:
        if not p:
            continue
        if be.name == "ref":
            be.lib.deviceFree(p)
        elif be.name == "oracle":
            C.CDLL(None).free(C.c_void_p(p))
        else:
            be.lib.DeviceFree(p, be.device)


def _read_raw(be, ptr, nbytes):
    """Reads engine-allocated memory (HyperLogLog allocates its outputs itself, hll.cu:117,150)."""
    import ctypes as C
    if not be.is_gpu:
        return np.frombuffer(C.string_at(ptr, nbytes), dtype=np.uint8).copy()
    host = np.zeros(max(nbytes, 1), np.uint8)
    be.lib.AsyncCopyDeviceToHost(host.ctypes.data, ptr, nbytes, be.space.stream, be.device)
    be.lib.WaitForCudaStream(be.space.stream, be.device)
    return host[:nbytes]


def run_hll(be, prev_dim, cur_values, nd, capacity, prev_size, batch, last=True, prev_values=None,
            prev_hash=None):
    import ctypes as C
    pd = be.put(np.asarray(prev_dim, np.uint8))
    pv = be.put(np.zeros(capacity, np.uint32) if prev_values is None else np.asarray(prev_values, np.uint32))
    ph = be.put(np.zeros(capacity, np.uint64) if prev_hash is None else np.asarray(prev_hash, np.uint64))
    pi = be.put(np.arange(capacity, dtype=np.uint32))
    cd = be.zeros(len(prev_dim))
    cv = be.put(np.asarray(cur_values, np.uint32))
    ch = be.zeros(8 * capacity)
    ci = be.put(np.arange(prev_size, prev_size + capacity, dtype=np.uint32))
    hll, size, cnt = C.c_void_p(), C.c_size_t(), C.c_void_p()
    n = be.lib.HyperLogLog(A.make_dimension_vector(pd.ptr, ph.ptr, pi.ptr, nd, capacity),
                           A.make_dimension_vector(cd.ptr, ch.ptr, ci.ptr, nd, capacity), pv.ptr, cv.ptr,
                           prev_size, batch, last, C.byref(hll), C.byref(size), C.byref(cnt),
                           be.space.stream, be.device)
    out = dict(n=n, dims=cd.get(np.uint8), hash=ch.get(np.uint64), index=ci.get(np.uint32), values=cv.get(np.uint32))
    if last and n > 0:
        out["hll"] = _read_raw(be, hll.value, size.value)
        out["counts"] = _read_raw(be, cnt.value, 2 * n).view(np.uint16)
        _free_outputs(be, hll.value, cnt.value)
    return out


def test_hll_sparse_mode(backend):
    """HyperLogLogTest.CheckSparseMode :1229 — byte-exact sparse register vector."""
    r = run_hll(backend, [1, 1, 2, 2, 3, 3, 4, 4] + [1] * 8,
                [0x010001, 0x020002, 0x010002, 0x020002, 0x010003, 0x020003, 0x010004, 0x020004],
                (0, 0, 0, 0, 1), 8, 0, 8)
    assert r["n"] == 4
    assert r["dims"].tolist() == [2, 4, 3, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0]
    assert r["hll"].tolist() == [2, 0, 3, 0, 4, 0, 3, 0, 3, 0, 3, 0, 1, 0, 2, 0, 2, 0, 3, 0]
    assert r["counts"].tolist() == [1, 1, 1, 2]


def test_hll_dense_mode(backend):
    """HyperLogLogTest.CheckDenseMode :1305 — a dim with 4996 registers goes dense (16384 B)."""
    prev = np.zeros(10000, np.uint8)
    prev[0:4] = [1, 1, 2, 2]
    prev[5000:] = 1
    vals = np.zeros(5000, np.uint32)
    vals[0:4] = [0x010001, 0x020002, 0x010002, 0x020002]
    vals[4:] = 0x010000 | np.arange(4996, dtype=np.uint32)
    r = run_hll(backend, prev, vals, (0, 0, 0, 0, 1), 5000, 0, 5000)
    assert r["n"] == 3
    exp_dims = np.zeros(10000, np.uint8)
    exp_dims[0:3] = [2, 0, 1]
    exp_dims[5000:5003] = 1
    assert r["dims"].tolist() == exp_dims.tolist()
    exp = np.zeros(16396, np.uint8)
    exp[0:4] = [2, 0, 3, 0]
    exp[4:5000] = 2
    exp[16388:16396] = [1, 0, 2, 0, 2, 0, 3, 0]
    assert r["hll"].size == 16396 and r["hll"].tolist() == exp.tolist()
    assert r["counts"].tolist() == [1, 4996, 2]
