"""Sizes either side of the single-launch finalize (one cluster of 8 CTAs, up to 32,768 groups) and of the dense HLL
directory (register arrays for 4,096 groups): exact group counts by construction, results against the oracle's legacy call
sequence (query/sort_reduce.cu order: hash ascending; query/hash_reduction.cu identity)."""
import numpy as np
import pytest

import harness as H
import hashes as HS
import test_pipeline_parity as T
from aresdb_b200 import cabi as A, expr as E, synth
from aresdb_b200.query import AggQuery, Measure

pytestmark = pytest.mark.gpu
TS, CITY, STATUS, FARE = T.TS, T.CITY, T.STATUS, T.FARE


def batch_with_groups(n: int, rows: int, seed: int = 7) -> synth.HostBatch:
    """`rows` rows over exactly `n` distinct request_at values (every value occurs), all other columns constant-ish."""
    rng = np.random.default_rng(seed)
    ts = synth.BASE_TS + np.concatenate([np.arange(n), rng.integers(0, n, rows - n)]).astype(np.uint32)
    rng.shuffle(ts)
    city = np.full(rows, 3, np.uint16)
    status = np.ones(rows, np.uint8)
    fare = (rng.integers(0, 6400, rows) / 64.0).astype(np.float32)
    ones = np.ones(rows, np.uint8)
    return synth.HostBatch([ts, city, status, fare], [ones, ones.copy(), ones.copy(), ones.copy()], rows, 0)


@pytest.mark.parametrize("mode", ["sort", "hash"])
@pytest.mark.parametrize("n", [1, 2, 1023, 8193, 32767, 32768, 32769, 40000])
def test_group_counts_around_the_single_launch_finalize(n, mode):
    eng, orc = H.get_backend("b200"), H.get_backend("oracle")
    hbs = [batch_with_groups(n, max(2 * n, 5000))]
    q = AggQuery([], [TS, CITY], Measure("sum", FARE), reduce_mode=A.ARES_REDUCE_HASH if mode == "hash" else A.ARES_REDUCE_SORT)
    exp = T.run_legacy(orc, q, hbs)
    for eg in (0, 50000):   # 0: the executor offers a 32768-row output first; 50000: group count asked first
        got = T.run_fused(eng, q, hbs, expected_groups=eg)
        if mode == "sort":
            assert got.groups == exp.groups == n
            T.assert_same_result(got, exp, ctx=f"n={n} eg={eg}")
        else:   # 32-bit identity: colliding rows merge; compare by hash
            by_hash = lambda r: dict(zip(HS.murmur3_32(r.packed_rows()).tolist(), r.measures.tolist()))
            assert got.groups == exp.groups and by_hash(got) == by_hash(exp), f"n={n} eg={eg}"


def test_dense_hll_directory_limit_is_reported():
    """4,096 groups fit the dense register arrays (byte-exact against the oracle's per-batch HyperLogLog sequence); one more
    is an error that names the remedy — never a wrong result — and the entry mode it names handles it."""
    import test_hll_pipeline as HP
    eng = H.get_backend("b200")
    q = AggQuery([], [TS], Measure("countdistincthll", CITY))
    full = [batch_with_groups(4096, 20000)]
    HP.assert_same_hll(HP.run_hll_fused(eng, q, full), HP._device_oracle(q, full), "4096 groups, dense")
    over = [batch_with_groups(4097, 20000)]
    with pytest.raises(A.AresError, match="dense HLL state: more than 4096"):
        HP.run_hll_fused(eng, q, over)
    HP.assert_same_hll(HP.run_hll_fused(eng, q, over, HP.ENTRY_MODE), HP._device_oracle(q, over), "4097 groups, entries")
