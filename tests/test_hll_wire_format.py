"""The application/hll wire format (aresdb_b200/hll_data.py) against the reference's own golden buffers
(testing/data/query/hll, hll_query_results, hll_empty_results — copied by tests/golden/make_hll_wire_fixture.py) and the
known answers of query/common/hll_test.go and query/hll_test.go; then from a real hll query of this repo's call sequence."""
import struct
from pathlib import Path

import numpy as np
import pytest

import harness as H
import test_hll_pipeline as HP
from aresdb_b200 import cabi as A, hll_data as W, synth
from aresdb_b200.postprocess import hll_nested_result

Z = np.load(Path(__file__).resolve().parent / "golden" / "hll_wire_format.npz")
DENSE = 16384


def _reference_inputs():
    """The inputs of the reference's SerializeHLL test (query/hll_test.go:27-82), its capacity-6 block compacted to 3 rows:
    dimensions Uint32 / enum byte / Int16, counts sparse-3, dense, sparse-4."""
    block = bytes([0, 0, 0, 0, 1, 0, 0, 0, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 2, 0, 2, 2, 0, 2, 3, 0, 1, 1, 0, 1, 1, 0, 1, 1])
    hll = bytearray(DENSE + 28)
    struct.pack_into("<3I", hll, 0, 0x00FF0001, 0x00FE0002, 0x00FD0003)
    hll[12] = hll[13] = 1
    struct.pack_into("<4I", hll, 12 + DENSE, 0x000100FF, 0x000200FE, 0x000300FD, 0x000400FC)
    return block, [3, DENSE, 4], bytes(hll)


EXPECTED = {"NULL": {"NULL": {"NULL": (3, [(1, 255), (2, 254), (3, 253)])}},
            "1": {"c": {"2": (2, "dense")}},
            "4294967295": {"d": {"514": (4, [(255, 1), (254, 2), (253, 3), (252, 4)])}}}


def _shape(res):
    def leaf(h):
        return (h.non_zero_registers, "dense" if h.dense is not None else h.sparse)
    return {a: {b: {c: leaf(h) for c, h in v2.items()} for b, v2 in v1.items()} for a, v1 in res.items()}


def test_sizes_known_answers():
    """CalculateSizes / CalculateEnumCasesBytes (query/common/hll_test.go:31-57)."""
    seven = list(range(7))
    assert W.header_and_total_size([0] * 5, seven, seven, {}, 100, 10, DENSE + 32) == (56, 16596)
    assert W.header_and_total_size([0] * 5, seven, seven, {1: ["a", "b", "c", "d"], 2: []}, 100, 10, DENSE + 32) == (88, 16628)
    assert [W.enum_cases_bytes(c) for c in (["ss", "a", "b"], ["ss"], [])] == [16, 8, 0]


def test_parser_reads_the_references_buffers():
    """NewTimeSeriesHLLResult / ParseHLLQueryResults / the empty result (query/common/hll_test.go:103-181)."""
    res = W.parse_hll_data(Z["hll"].tobytes())
    assert _shape(res) == EXPECTED
    dense = res["1"]["c"]["2"].dense
    assert len(dense) == DENSE and dense[0] == 1 and dense[1] == 1 and not any(dense[2:])
    results, errors = W.parse_hll_query_results(Z["hll_query_results"].tobytes())
    assert errors == [None, "test"] and results[1] is None and _shape(results[0]) == EXPECTED
    assert W.parse_hll_query_results(Z["hll_empty_results"].tobytes()) == ([{}], [None])
    # ignoreEnum: enum dimensions stay numbers
    assert set(W.parse_hll_data(Z["hll"].tobytes(), ignore_enum=True)["1"]) == {"2"}
    with pytest.raises(ValueError):
        W.parse_hll_query_results(b"\x00" * 16)


def test_serializer_reproduces_the_references_buffers_byte_for_byte():
    block, counts, hll = _reference_inputs()
    types = [W.MEM_UINT32, W.MEM_UINT8, W.MEM_INT16]
    payload = W.serialize_hll_data([0, 0, 1, 1, 1], [0, 2, 1], types, {1: ["a", "b", "c", "d"]}, block, counts, hll)
    assert payload == Z["hll"].tobytes()
    w = W.HLLQueryResultsWriter()
    w.write_result(payload)
    w.write_error("test")
    assert w.get_bytes() == Z["hll_query_results"].tobytes()
    e = W.HLLQueryResultsWriter()
    e.write_result(b"")
    assert e.get_bytes() == Z["hll_empty_results"].tobytes()
    # the slices query/hll_test.go:83-103 checks (enum dictionary a, b, c: header 64 bytes)
    data = W.serialize_hll_data([0, 0, 1, 1, 1], [0, 2, 1], types, {1: ["a", "b", "c"]}, block, counts, hll)
    assert len(data) == 104 + DENSE + 32 and data[64:94] == block
    assert data[96:102] == np.array(counts, np.uint16).tobytes() and data[104:104 + DENSE + 28] == hll


@pytest.mark.parametrize("name", ["sparse_by_city", "dense_by_status", "two_dims"])
@pytest.mark.parametrize("backend", ["ref", "oracle"])
def test_a_real_hll_result_round_trips_through_the_wire_format(backend, name):
    """An hll query through the reference call sequence (HOST build / C restatement) -> HLLResult -> payload -> parser:
    same groups, same register sets; enum names and a time-zone shift applied on the way."""
    be = H.get_backend(backend)
    if backend == "oracle":
        HP.set_oracle_device_semantics(False)
    q = HP.hll_queries()[name]
    hbs = [synth.generate_batch(d, n, num_cities=7, null_rate=0.03) for d, n in ((0, 9000), (1, 6000))]
    res = HP.run_hll_query(be, q, hbs)
    mem = {A.Uint32: W.MEM_UINT32, A.Uint16: W.MEM_UINT16, A.Uint8: W.MEM_UINT8, A.Int32: W.MEM_INT32, A.Int16: W.MEM_INT16,
           A.Int8: W.MEM_INT8, A.Bool: W.MEM_BOOL, A.Float32: W.MEM_FLOAT32}
    types = [mem[t] for t in q.dim_types]
    payload = W.serialize_hll_result(res, types)
    parsed = W.parse_hll_data(payload)
    expect = hll_nested_result(res)          # {dim strings ...: estimate} of the same result

    def walk(a, b, depth):
        assert set(a) == set(b)
        for k in a:
            if depth == len(q.dimensions) - 1:
                assert isinstance(a[k], W.HLL)
            else:
                walk(a[k], b[k], depth + 1)
    walk(parsed, expect, 0)
    # register sets: every group's dense registers equal what the result holds
    dense = res.dense_registers()
    rows = res.dims.rows
    cols = res.dims.decoded_dims()
    for g in range(res.groups):
        cur = parsed
        for d in range(len(q.dimensions)):
            v = cols[d][g]
            cur = cur["NULL" if v is None else str(int(v))]
        assert (cur.dense_registers() == dense[rows[g]]).all()
        assert cur.non_zero_registers == int(res.counts[g])
    if name == "two_dims":
        # a query that ran in a time zone: the day bucket (dimension 0) goes back to an instant, clamped at 0 (SerializeHLL's
        # "fix time dimension" step with utils.AdjustOffset)
        shifted = W.parse_hll_data(W.serialize_hll_result(res, types, time_dimensions=[0], from_offset=-28800, to_offset=-28800))
        days = sorted(int(k) for k in parsed if k != "NULL")
        assert sorted(int(k) for k in shifted if k != "NULL") == [d + 28800 for d in days]
        assert ("NULL" in shifted) == ("NULL" in parsed)


def test_register_set_operations_known_answers():
    """Set / Merge / ConvertToDense / ConvertToSparse / Encode / EncodeBinary / Decode / Compute
    (query/common/hll_test.go:160-262)."""
    h = W.HLL(2, sparse=[(100, 1), (200, 2)])
    assert h.compute() == 2.0
    back = W.HLL.decode(h.encode())
    assert (back.non_zero_registers, back.sparse, back.dense) == (2, [(100, 1), (200, 2)], None)
    dense = bytearray(DENSE)
    dense[100], dense[200] = 1, 2
    d = W.HLL(2, dense=bytes(dense))
    back = W.HLL.decode(d.encode())
    assert (back.non_zero_registers, back.sparse, back.dense) == (2, None, bytes(dense))
    one = W.HLL(1, sparse=[(100, 1)])
    assert one.encode_binary() == bytes([100, 0, 1, 0])
    hll, off = W.read_hll(one.encode_binary(), 1, 0)
    assert (hll.non_zero_registers, hll.sparse, off) == (1, [(100, 1)], 4)
    assert W.HLL(1, sparse=[(1, 255)]).encode_binary() == bytes([1, 0, 255, 255])      # int8(rho) sign-extends
    s = W.HLL(0)
    s.set(100, 1)
    s.set(200, 2)
    assert (s.non_zero_registers, s.sparse, s.dense) == (2, [(100, 1), (200, 2)], None)
    for i in range(201, 4300):
        s.set(i, 3)
    assert s.sparse is None and len(s.dense) == DENSE and s.non_zero_registers == 4101
    assert [s.dense[i] for i in (100, 200, 201, 4299, 4300)] == [1, 2, 3, 3, 0]
    assert not s.to_sparse()                                                              # more than a quarter of the registers
    a, b = W.HLL(2, sparse=[(5, 3), (7, 1)]), W.HLL(2, sparse=[(5, 2), (9, 4)])
    a.merge(b)
    assert a.non_zero_registers == 3 and [a.dense[i] for i in (5, 7, 9)] == [3, 1, 4]
    assert a.to_sparse() and sorted(a.sparse) == [(5, 3), (7, 1), (9, 4)] and a.dense is None


def test_build_vectors_from_a_nested_result_known_answer():
    """BuildVectorsFromHLLResult (query/common/hll_test.go:264-318): keys in string order, children first, enum names
    back to ids, small register sets sparse."""
    dense = bytearray(DENSE)
    dense[0] = dense[1] = 1
    two = lambda: W.HLL(2, dense=bytes(dense))
    result = {"NULL": {"NULL": {"NULL": W.HLL(3, sparse=[(1, 255), (2, 254), (3, 253)])}},
              "1": {"c": {"2": two(), "3": two()}},
              "4294967295": {"d": {"514": W.HLL(4, sparse=[(255, 1), (254, 2), (253, 3), (252, 4)])}, "e": {"4": two()}}}
    hll, dims, counts = W.build_vectors_from_hll_result(result, [W.MEM_UINT32, W.MEM_UINT8, W.MEM_INT16], {1: {"c": 0, "d": 1, "e": 2}}, [0, 2, 1])
    assert hll == bytes([0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 1, 0, 1, 0, 1, 0, 255, 0, 1, 0, 254, 0, 2, 0, 253, 0, 3, 0, 252, 0, 4, 0,
                         0, 0, 1, 0, 1, 0, 1, 0, 1, 0, 255, 255, 2, 0, 254, 255, 3, 0, 253, 255])
    assert dims == bytes([1, 0, 0, 0, 1, 0, 0, 0, 255, 255, 255, 255, 255, 255, 255, 255, 0, 0, 0, 0,
                          2, 0, 3, 0, 2, 2, 4, 0, 0, 0,
                          0, 0, 1, 2, 0,
                          1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0])
    assert counts == bytes([2, 0, 2, 0, 4, 0, 2, 0, 3, 0])
    # and the vectors serialize into a payload the parser reads back to the same nested result
    payload = W.serialize_hll_data([0, 0, 1, 1, 1], [0, 2, 1], [W.MEM_UINT32, W.MEM_UINT8, W.MEM_INT16], {1: ["c", "d", "e"]},
                                   dims, np.frombuffer(counts, np.uint16), hll)
    back = W.parse_hll_data(payload)
    assert set(back) == set(result) and set(back["4294967295"]) == {"d", "e"} and set(back["1"]["c"]) == {"2", "3"}
    assert back["NULL"]["NULL"]["NULL"].sparse == [(1, 255), (2, 254), (3, 253)]
