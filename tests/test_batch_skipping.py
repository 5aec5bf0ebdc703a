"""Zone-map batch skipping (aresdb_b200/skipping.py), the reference's shouldSkipLiveBatchWithFilter
(query/aql_processor.go:1447-1526) on the same numbers the engine gets as BatchPlan.Ranges."""
import numpy as np
import pytest

import harness as H
from aresdb_b200 import cabi as A
from aresdb_b200 import expr as E, synth
from aresdb_b200.query import AggQuery, Measure
from aresdb_b200.skipping import filter_excludes_range, should_skip_batch
import test_pipeline_parity as T

TS, CITY, STATUS, FARE = T.TS, T.CITY, T.STATUS, T.FARE
R = {synth.COL_REQUEST_AT: (1000, 2000), synth.COL_CITY_ID: (5, 9)}


@pytest.mark.parametrize("make, lit, skip", [
    (E.ge, 2001, True), (E.ge, 2000, False), (E.gt, 2000, True), (E.gt, 1999, False),
    (E.le, 999, True), (E.le, 1000, False), (E.lt, 1000, True), (E.lt, 1001, False),
    (E.eq, 999, True), (E.eq, 2001, True), (E.eq, 1500, False), (E.ne, 5000, False)])
def test_operators(make, lit, skip):
    assert filter_excludes_range(make(TS, E.Lit(lit)), R) == skip


def test_literal_on_the_left_mirrors_the_operator():
    assert filter_excludes_range(E.le(E.Lit(2001), TS), R)          # 2001 <= ts  ==  ts >= 2001
    assert not filter_excludes_range(E.le(E.Lit(2000), TS), R)
    assert filter_excludes_range(E.gt(E.Lit(1000), TS), R)          # 1000 > ts   ==  ts < 1000
    assert not filter_excludes_range(E.gt(E.Lit(1001), TS), R)


def test_only_plain_column_vs_integer_literal_filters_qualify():
    assert not filter_excludes_range(E.ge(E.floor(TS, E.Lit(10)), E.Lit(5000)), R)     # expression, not a column
    assert not filter_excludes_range(E.ge(FARE, E.Lit(1e9)), R)                          # float column, no entry
    assert not filter_excludes_range(E.ge(STATUS, E.Lit(200)), R)                        # no zone-map entry
    assert not filter_excludes_range(E.or_(E.ge(TS, E.Lit(5000)), E.eq(CITY, E.Lit(7))), R)
    assert not filter_excludes_range(E.ge(TS, E.Lit(5000.0)), R)                         # float literal


def test_query_level_and_no_zone_map():
    q = AggQuery([E.eq(STATUS, E.Lit(1)), E.ge(TS, E.Lit(5000))], [CITY], Measure("count"))
    assert should_skip_batch(q, R) and not should_skip_batch(q, None) and not should_skip_batch(q, {})
    assert not should_skip_batch(AggQuery([E.eq(STATUS, E.Lit(1))], [CITY], Measure("count")), R)


def test_skipping_never_changes_a_result():
    """Reference call sequence on the C restatement: dropping the batches should_skip_batch names gives the same groups."""
    orc = H.get_backend("oracle")
    hbs = [synth.generate_batch(d, 3000, num_cities=10) for d in range(4)]
    t0 = synth.BASE_TS
    q = AggQuery([E.ge(TS, E.Lit(t0 + 86400 + 100)), E.lt(TS, E.Lit(t0 + 3 * 86400))], [E.floor(TS, E.Lit(3600)), CITY],
                 Measure("sum", FARE))
    keep = [hb for hb in hbs if not should_skip_batch(q, synth.zone_map(hb))]
    assert len(keep) == 2                                                 # days 1 and 2 survive, 0 and 3 are skipped
    T.assert_same_result(T.run_legacy(orc, q, keep), T.run_legacy(orc, q, hbs), ctx="skipped batches")


@pytest.mark.gpu
def test_fused_executor_skips_batches():
    eng, orc = H.get_backend("b200"), H.get_backend("oracle")
    hbs = [synth.generate_batch(d, 20000, num_cities=10) for d in range(3)]
    t0 = synth.BASE_TS
    q = AggQuery([E.ge(TS, E.Lit(t0 + 86400)), E.eq(STATUS, E.Lit(1))], [E.floor(TS, E.Lit(3600)), CITY], Measure("count"))
    from aresdb_b200.executor import FusedBatchExecutor
    ex = FusedBatchExecutor(eng.lib, eng.space, q)
    keep = []
    for hb in hbs:
        b = T.upload(eng, hb, ranges=synth.zone_map(hb))
        keep.append(b)
        ex.process_batch(b)
    got = ex.result()
    ex.close()
    assert ex.skipped == 1 and ex.calls == 2
    T.assert_same_result(got, T.run_legacy(orc, q, hbs), ctx="fused + skipping")


def test_values_that_wrap_in_int32_are_never_skipped_on():
    """A uint32 value >= 2^31 compares as a negative int32 against an integer literal (the reference's promotion); the
    zone map's unbounded integers would say otherwise, so such ranges / literals never skip."""
    from aresdb_b200 import cabi as A, expr as E
    from aresdb_b200.skipping import filter_excludes_range
    col = E.Col(0, A.Uint32)
    big = {0: (2 ** 31 + 5, 2 ** 31 + 9)}
    assert not filter_excludes_range(E.resolve(E.lt(col, E.Lit(100))), big)     # int32: negative < 100 is TRUE for every row
    assert not filter_excludes_range(E.resolve(E.gt(col, E.Lit(2 ** 31 + 20))), {0: (1, 2)})
    assert filter_excludes_range(E.resolve(E.lt(col, E.Lit(100))), {0: (200, 300)})
    assert filter_excludes_range(E.resolve(E.ge(col, E.Lit(-5))), {0: (200, 300)}) is False
