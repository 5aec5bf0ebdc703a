"""BASELINE-size checks of the fused path against an INDEPENDENT implementation (plain torch ops on the
same device-resident columns: boolean masks + index_add in float64), plus size-independent properties.
cfg2: 1e8 rows, 1 filter + SUM group-by 1 dim.  cfg3 shape at 2.5e7 rows per batch x 2 batches."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _columns(bufs, values_off, rows):
    import torch
    ts = bufs[0][values_off:values_off + 4 * rows].view(torch.int32)
    city = bufs[1][values_off:values_off + 2 * rows].view(torch.int16)
    status = bufs[2][values_off:values_off + rows]
    fare = bufs[3][values_off:values_off + 4 * rows].view(torch.float32)
    valid = []
    for b in bufs:
        bits = b[: (rows + 7) // 8]
        v = ((bits.unsqueeze(1) >> torch.arange(8, device=b.device, dtype=torch.uint8)) & 1).reshape(-1)[:rows].bool()
        valid.append(v)
    return (ts, city, status, fare), valid


def test_cfg2_1e8_rows_sum_by_city():
    import torch
    import harness as H
    from aresdb_b200 import cabi as A, columns, expr as E, synth
    from aresdb_b200.executor import Batch, FusedBatchExecutor
    from aresdb_b200.query import AggQuery, Measure
    eng = H.get_backend("b200")
    rows = 100_000_000
    dev = torch.device("cuda:0")
    bufs, voff = synth.generate_batch_cuda(0, rows, dev, num_cities=100)
    cols = [columns.slice_of(b.data_ptr(), dt, rows, 0, voff, 2) for b, dt in zip(bufs, synth.COLUMN_TYPES)]
    TS, CITY, STATUS, FARE = (E.Col(i, t) for i, t in enumerate(synth.COLUMN_TYPES))
    q = AggQuery([E.eq(STATUS, E.Lit(1))], [CITY], Measure("sum", FARE))
    ex = FusedBatchExecutor(eng.lib, eng.space, q)
    ex.process_batch(Batch(cols, rows))
    res = ex.result()
    ex.close()
    (ts, city, status, fare), (vts, vcity, vstatus, vfare) = _columns(bufs, voff, rows)
    keep = vstatus & (status == 1)
    # group = (city value, city validity); NULL fare contributes the identity 0
    gid = city.to(torch.int64) * 2 + vcity.to(torch.int64)
    contrib = torch.where(vfare, fare.double(), torch.zeros((), dtype=torch.float64, device=dev))
    sums = torch.zeros(2 * 65536, dtype=torch.float64, device=dev).index_add_(0, gid[keep], contrib[keep])
    present = torch.zeros(2 * 65536, dtype=torch.bool, device=dev)
    present[gid[keep]] = True
    exp = {}
    for g in torch.nonzero(present).flatten().tolist():
        c, v = g // 2, g % 2
        exp[np.uint16(c).tobytes() + bytes([v])] = sums[g].item()
    got = res.as_dict()
    assert got.keys() == exp.keys()
    # fares are multiples of 1/64: every partial sum is exactly representable -> bit-exact in any order
    assert all(got[k] == exp[k] for k in exp)
    # properties: number of groups, total of the sums, hash-ascending order of the output
    assert res.groups == 101  # 100 cities + the NULL-city group
    assert abs(sum(got.values()) - contrib[keep].sum().item()) < 1e-3


def test_cfg3_shape_counts_match_torch():
    import torch
    import harness as H
    from aresdb_b200 import columns, expr as E, synth
    from aresdb_b200.executor import Batch, FusedBatchExecutor
    from aresdb_b200.query import AggQuery, Measure
    eng = H.get_backend("b200")
    rows = 25_000_000
    dev = torch.device("cuda:0")
    TS, CITY, STATUS, FARE = (E.Col(i, t) for i, t in enumerate(synth.COLUMN_TYPES))
    q = AggQuery([E.eq(STATUS, E.Lit(1)), E.gt(FARE, E.Lit(5.0)), E.ne(CITY, E.Lit(0))],
                 [E.floor(TS, E.Lit(3600)), CITY], Measure("count"))
    ex = FusedBatchExecutor(eng.lib, eng.space, q)
    total = 0
    keep_alive = []
    exp_counts = {}
    for day in range(2):
        bufs, voff = synth.generate_batch_cuda(day, rows, dev, num_cities=100)
        keep_alive.append(bufs)
        cols = [columns.slice_of(b.data_ptr(), dt, rows, 0, voff, 2) for b, dt in zip(bufs, synth.COLUMN_TYPES)]
        ex.process_batch(Batch(cols, rows))
        (ts, city, status, fare), (vts, vcity, vstatus, vfare) = _columns(bufs, voff, rows)
        keep = vstatus & (status == 1) & vfare & (fare > 5.0) & vcity & (city != 0)
        total += int(keep.sum().item())
    res = ex.result()
    ex.close()
    assert int(res.measures.astype(np.int64).sum()) == total       # every surviving row is counted exactly once
    hashes_sorted = res.rows                                        # group order is the reference's hash order:
    assert len(set(hashes_sorted)) == res.groups                    # ... and groups are distinct
    assert res.groups <= 2 * 25 * 101 * 2
