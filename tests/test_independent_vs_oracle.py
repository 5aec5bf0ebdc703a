"""Pins tests/independent.py (the plain-torch restatement used for the BASELINE-size GPU checks and by
bench.py's self-verification) against the oracle at sizes the oracle finishes in seconds."""
import numpy as np
import pytest

import harness as H
import independent as I
import test_pipeline_parity as T
from aresdb_b200 import cabi as A, expr as E, synth
from aresdb_b200.query import AggQuery, Measure

TS, CITY, STATUS, FARE = T.TS, T.CITY, T.STATUS, T.FARE
DAYS = 3


def _queries():
    t0 = synth.BASE_TS
    return {
        "cfg3": AggQuery([E.eq(STATUS, E.Lit(1)), E.gt(FARE, E.Lit(5.0)), E.ne(CITY, E.Lit(0)),
                          E.ge(TS, E.Lit(t0 + 1800)), E.lt(TS, E.Lit(t0 + DAYS * 86400 - 1800))],
                         [E.floor(TS, E.Lit(3600)), CITY], Measure("sum", FARE)),
        "cfg3_count": AggQuery([E.eq(STATUS, E.Lit(1)), E.gt(FARE, E.Lit(5.0)), E.ne(CITY, E.Lit(0))],
                               [E.floor(TS, E.Lit(3600)), CITY], Measure("count")),
        "cfg2": AggQuery([E.eq(STATUS, E.Lit(1))], [CITY], Measure("sum", FARE)),
        "cfg4": AggQuery([], [CITY, E.floor(TS, E.Lit(60))], Measure("sum", FARE), reduce_mode=A.ARES_REDUCE_HASH),
        "cfg4_hll": AggQuery([E.eq(STATUS, E.Lit(1))], [E.floor(TS, E.Lit(86400)), CITY], Measure("countdistincthll", TS)),
    }


@pytest.fixture(scope="module")
def host_batches():
    return [synth.generate_batch(d, n, num_cities=40, null_rate=0.03) for d, n in ((0, 30000), (1, 20011), (2, 8))]


def _expected(name, hbs):
    t0 = synth.BASE_TS
    exp = I.Expected(name, DAYS, "cpu", t0, t0 + 1800, t0 + DAYS * 86400 - 1800)
    for hb in hbs:
        bufs, voff = I.host_batch_buffers(hb)
        exp.add_batch(bufs, voff, hb.num_rows, chunk=1 << 14)
    return exp


@pytest.mark.parametrize("name", ["cfg3", "cfg3_count", "cfg2", "cfg4"])
def test_independent_matches_oracle(name, host_batches):
    orc = H.get_backend("oracle")
    res = T.run_legacy(orc, _queries()[name], host_batches)
    out = _expected(name, host_batches).check(res)
    assert out["groups"] == res.groups > 0


def test_independent_hll_matches_oracle(host_batches):
    orc = H.get_backend("oracle")
    from aresdb_b200.executor import LegacyBatchExecutor
    q = _queries()["cfg4_hll"]
    ex = LegacyBatchExecutor(orc.lib, orc.space, q)
    for i, hb in enumerate(host_batches):
        ex.process_batch(T.upload(orc, hb), is_last=i == len(host_batches) - 1)
    out = _expected("cfg4_hll", host_batches).check_hll(ex.hll)
    assert out["groups"] == ex.hll.groups > 0


def test_independent_detects_a_wrong_sum(host_batches):
    orc = H.get_backend("oracle")
    res = T.run_legacy(orc, _queries()["cfg2"], host_batches)
    res.measures[3] = np.nextafter(res.measures[3], np.inf)
    with pytest.raises(AssertionError):
        _expected("cfg2", host_batches).check(res)
