"""Result post-processing rules (aresdb_b200/postprocess.py) against hand-derived expectations from the
reference's formatting code (query/common/dimval.go:36-212, query/aql_postprocessor.go:232-264,
query/common/hll.go:735-775)."""
import math

import numpy as np
import pytest

from aresdb_b200 import cabi as A, expr as E, postprocess as PP
from aresdb_b200.postprocess import DimensionMeta as M
from aresdb_b200.query import AggQuery, HLLResult, Measure, QueryResult

TS = 1700012345   # 2023-11-15 01:39:05 UTC, a Wednesday


@pytest.mark.parametrize("x,s", [(1.5, "1.5"), (100.0, "100"), (0.1, "0.1"), (1e6, "1e+06"), (123456.0, "123456"),
                                 (1.0e-5, "1e-05"), (0.0001, "0.0001"), (3.4028235e38, "3.4028235e+38"), (0.0, "0"),
                                 (86.074904, "86.074905"), (-2.5, "-2.5"), (1e21, "1e+21")])
def test_float32_formatting_matches_go_shortest_g(x, s):
    assert PP.format_float32(x) == s


@pytest.mark.parametrize("bucketizer,val,exp", [
    ("hour", TS - TS % 3600, "2023-11-15 01:00"),
    ("day", TS - TS % 86400, "2023-11-15"),
    ("minute", TS - TS % 60, "2023-11-15 01:39"),
    ("15m", TS - TS % 900, "2023-11-15 01:30"),
    ("4 hours", TS - TS % 14400, "2023-11-15 00:00"),
    ("time of day", TS % 86400, "01:39"),
    ("hour of day", (TS % 86400) // 3600 * 3600, "01:00"),
    ("hour of week", ((TS - 345600) % 604800) // 3600 * 3600, "Wednesday 01:00"),
    ("day of week", ((TS - 345600) % 604800) // 86400, "Wednesday"),
    ("month", 1698796800, "1698796800"),           # irregular buckets stay epoch seconds
])
def test_time_dimension_formatting(bucketizer, val, exp):
    assert PP.format_time_dimension(val, M(time_bucketizer=bucketizer)) == exp


def test_time_unit_overrides_bucketizer():
    assert PP.format_time_dimension(7200, M(time_bucketizer="hour", time_unit="hour")) == "2"
    assert PP.format_time_dimension(7200, M(time_bucketizer="hour", time_unit="millisecond")) == "7200000"
    assert PP.format_time_dimension(7200, M(time_bucketizer="hour", time_unit="second")) == "7200"


def _result(q, dims, valid, measures):
    """Builds the binary block of a QueryResult (capacity == groups) from per-dimension arrays."""
    g = len(measures)
    parts = []
    for p, qi in enumerate(q.dim_order):
        parts.append(np.ascontiguousarray(dims[qi]).view(np.uint8).reshape(-1))
    for p, qi in enumerate(q.dim_order):
        parts.append(np.asarray(valid[qi], np.uint8))
    block = np.concatenate(parts)
    return QueryResult(q, block, g, np.ascontiguousarray(measures).view(np.uint8).reshape(-1), g)


def test_nested_result_with_enum_null_and_time_dims():
    ts, city, status = E.Col(0, A.Uint32, "ts"), E.Col(1, A.Uint16, "city"), E.Col(2, A.Uint8, "status")
    q = AggQuery([], [E.floor(ts, E.Lit(3600)), status, city], Measure("count"))
    hours = np.array([TS - TS % 3600, TS - TS % 3600, TS - TS % 3600 + 3600], np.uint32)
    res = _result(q, [hours, np.array([0, 1, 5], np.uint8), np.array([7, 7, 0], np.uint16)],
                  [[1, 1, 1], [1, 1, 1], [1, 1, 0]], np.array([5, 4, 9], np.uint32))
    out = PP.nested_result(res, [M(time_bucketizer="hour"), M(enum_names=["completed", "canceled"]), None])
    assert out == {"2023-11-15 01:00": {"completed": {"7": 5.0}, "canceled": {"7": 4.0}},
                   "2023-11-15 02:00": {"5": {"NULL": 9.0}}}     # id outside the dictionary prints as a number


def test_measure_types_become_float64():
    city, fare = E.Col(0, A.Uint16, "city"), E.Col(1, A.Float32, "fare")
    q = AggQuery([], [city], Measure("sum", fare))
    res = _result(q, [np.array([3], np.uint16)], [[1]], np.array([12.625], np.float64))
    assert PP.nested_result(res) == {"3": 12.625}
    q = AggQuery([], [city], Measure("max", E.Unary(A.Negate, city)))       # signed 4-byte measure
    res = _result(q, [np.array([3], np.uint16)], [[1]], np.array([-3], np.int32))
    assert PP.nested_result(res) == {"3": -3.0}


def test_hll_estimate_linear_counting_and_large_range():
    m = 16384.0
    dense = np.zeros(16384, np.uint8)
    assert PP.hll_estimate(dense) == 0.0
    dense[:100] = 1                               # 100 registers hit: linear counting m * ln(m / (m - 100))
    assert PP.hll_estimate(dense) == float(int(m * math.log(m / (m - 100))))
    dense[:] = 12                                 # all registers rho+1 = 12: raw estimate, far above 5m
    assert PP.hll_estimate(dense) == float(int(0.7213 / (1 + 1.079 / m) * m * m / (m * 2.0 ** -12)))


def test_hll_nested_result():
    city = E.Col(0, A.Uint16, "city")
    q = AggQuery([], [city], Measure("countdistincthll", E.Col(1, A.Uint32, "x")))
    block = np.concatenate([np.array([7, 9], np.uint16).view(np.uint8), np.array([1, 1], np.uint8)])
    regs = np.concatenate([np.array([(3 << 16) | 5, (1 << 16) | 9], np.uint32).view(np.uint8),     # group 7: 2 registers
                           np.array([(2 << 16) | 1], np.uint32).view(np.uint8)])                    # group 9: 1 register
    r = HLLResult(q, 2, block, 2, regs, np.array([2, 1], np.uint16))
    assert PP.hll_nested_result(r) == {"7": 2.0, "9": 1.0}


def test_hll_estimate_reference_known_answer_and_bias_range():
    """`Computes hll correctly` (reference query/common/hll_test.go:157-170): two registers -> 2.0; and the mid range
    (15.5k < n <= 5m) goes through the empirical bias table exactly as getEstimateBias does."""
    import bisect
    from aresdb_b200 import hll_bias_p14 as B
    dense = np.zeros(16384, np.uint8)
    dense[100], dense[200] = 1, 2
    assert PP.hll_estimate(dense) == 2.0
    # a register set whose raw estimate lies inside the table: every register hit once with rho+1 = 2
    dense[:] = 2
    m = 16384.0
    raw = 0.7213 / (1 + 1.079 / m) * m * m / sum(1.0 / 4.0 for _ in range(16384))
    assert B.RAW_ESTIMATES[0] < raw < B.RAW_ESTIMATES[-1] and raw <= 5 * m
    # independent restatement of the neighbour rule: 6 nearest by squared distance
    near = sorted(range(len(B.RAW_ESTIMATES)), key=lambda j: (B.RAW_ESTIMATES[j] - raw) ** 2)[:6]
    want = raw - sum(B.BIASES[j] for j in near) / 6.0
    assert PP.hll_estimate(dense) == float(int(want))
    assert abs(PP.hll_estimate_bias(raw) - sum(B.BIASES[j] for j in near) / 6.0) < 1e-9
    # table edges: below the first and beyond the last raw estimate the window shrinks but k stays 6
    assert PP.hll_estimate_bias(1000.0) == sum(B.BIASES[:6]) / 6.0
    assert PP.hll_estimate_bias(1e6) == sum(B.BIASES[-6:]) / 6.0


def test_merging_nested_results_of_several_nodes_known_answers():
    """broker/result_merge.go through its own cases (broker/result_merge_test.go:26-459): same and different shapes for
    sum / count / max / min, avg = sum result / count result with every key on both sides, hll = register-set merge."""
    import copy
    from aresdb_b200.postprocess import MergeError, merge_nested_results
    a = {"1234": {"foo": 123, "bar": 2}}
    b = {"1234": {"foo": 1, "bar": 1}}
    for agg, exp in (("sum", {"1234": {"foo": 124, "bar": 3}}), ("count", {"1234": {"foo": 124, "bar": 3}}),
                     ("max", {"1234": {"foo": 123, "bar": 2}}), ("min", {"1234": {"foo": 1, "bar": 1}})):
        assert merge_nested_results(copy.deepcopy(a), copy.deepcopy(b), agg) == exp
        assert merge_nested_results({}, {}, agg) == {}
        # different shapes: a key on one side only is taken as it is
        assert merge_nested_results({"1234": {"foo": 123}}, copy.deepcopy(b), agg) == {"1234": {"foo": exp["1234"]["foo"], "bar": 1}}
        assert merge_nested_results({}, copy.deepcopy(b), agg) == b
        assert merge_nested_results({"1234": {"foo": 123}}, {}, agg) == {"1234": {"foo": 123}}
    assert merge_nested_results({"1234": {"foo": 2, "bar": 1}}, {"1234": {"foo": 1, "bar": 2}}, "avg") == {"1234": {"foo": 2, "bar": 0.5}}
    assert merge_nested_results({}, {}, "avg") == {}
    for l, r in (({"1234": {"foo": 2}}, b), ({}, b), ({"1234": {"foo": 123}}, {})):
        with pytest.raises(MergeError, match="error calculating avg"):
            merge_nested_results(copy.deepcopy(l), copy.deepcopy(r), "avg")
    with pytest.raises(MergeError, match="different type"):
        merge_nested_results({"k": {"x": 1}}, {"k": 2}, "sum")
    # hll: a result merged with itself is itself (the reference's case on its golden buffer), and a real union
    from pathlib import Path
    from aresdb_b200 import hll_data as W
    z = np.load(Path(__file__).resolve().parent / "golden" / "hll_wire_format.npz")
    lhs = W.parse_hll_query_results(z["hll_query_results"].tobytes(), True)[0][0]
    rhs = W.parse_hll_query_results(z["hll_query_results"].tobytes(), True)[0][0]
    before = {k: v3.dense_registers().tobytes() for k, v1 in lhs.items() for v2 in v1.values() for v3 in v2.values()}
    merged = merge_nested_results(lhs, rhs, "hll")
    after = {k: v3.dense_registers().tobytes() for k, v1 in merged.items() for v2 in v1.values() for v3 in v2.values()}
    assert after == before and merged["NULL"]["NULL"]["NULL"].non_zero_registers == 3
    x, y = {"g": W.HLL(2, sparse=[(5, 3), (7, 1)])}, {"g": W.HLL(2, sparse=[(5, 2), (9, 4)]), "h": W.HLL(1, sparse=[(1, 1)])}
    m = merge_nested_results(x, y, "hll")
    assert m["g"].non_zero_registers == 3 and m["g"].compute() == pytest.approx(3.0, abs=0.01) and m["h"].sparse == [(1, 1)]
    with pytest.raises(MergeError, match="HLL value found"):
        merge_nested_results({"g": W.HLL(1, sparse=[(1, 1)])}, {"g": W.HLL(1, sparse=[(2, 1)])}, "sum")
