"""ComputeColumnRanges (the engine's own zone-map production) against numpy min / max of the valid values."""
import numpy as np
import pytest

import harness as H
from aresdb_b200 import cabi as A, columns, synth
from aresdb_b200.executor import compute_zone_map

pytestmark = pytest.mark.gpu


def _expect(dt, values, valid):
    sel = np.asarray(values)[np.asarray(valid) != 0] if valid is not None else np.asarray(values)
    if sel.size == 0:
        return None
    if dt == A.Float32:
        if not np.isfinite(sel).all() or np.signbit(sel).any():
            return None
        return int(sel.min().view(np.uint32)), int(sel.max().view(np.uint32))
    lo, hi = int(sel.min()), int(sel.max())
    return (lo, hi) if lo >= 0 and hi < 2 ** 31 else None


@pytest.mark.parametrize("start_bit", [0, 5])
@pytest.mark.parametrize("rows", [1, 31, 32, 1000, 100_003])
def test_ranges_match_numpy(rows, start_bit):
    eng = H.get_backend("b200")
    rng = np.random.default_rng(rows + start_bit)
    specs = [
        (A.Uint32, rng.integers(1_726_963_200, 1_726_963_200 + 86400, rows).astype(np.uint32), rng.random(rows) > 0.1),
        (A.Uint16, rng.integers(3, 900, rows).astype(np.uint16), rng.random(rows) > 0.5),
        (A.Uint8, rng.integers(0, 4, rows).astype(np.uint8), None),
        (A.Int16, rng.integers(-5, 300, rows).astype(np.int16), rng.random(rows) > 0.2),           # negative -> unknown (mostly)
        (A.Int32, rng.integers(7, 2_000_000, rows).astype(np.int32), rng.random(rows) > 0.2),
        (A.Float32, (rng.integers(0, 6400, rows) / 64.0).astype(np.float32), rng.random(rows) > 0.05),
        (A.Float32, (rng.integers(-10, 6400, rows) / 64.0).astype(np.float32), None),                # negative floats -> unknown
        (A.Bool, rng.integers(0, 2, rows).astype(np.uint8), rng.random(rows) > 0.3),
        (A.Uint32, np.zeros(rows, np.uint32), np.zeros(rows, bool)),                                 # no valid value
    ]
    cols, keep, want = [], [], {}
    for i, (dt, v, ok) in enumerate(specs):
        buf, vp = columns.make_column(eng.space, dt, v, valid=None if ok is None else ok.astype(np.uint8), start_bit=start_bit)
        cols.append(vp)
        keep.append(buf)
        e = _expect(dt, v, ok)
        if e is not None:
            want[i] = e
    cols.append(columns.constant_column(A.Uint16, 42, True))
    want[len(specs)] = (42, 42)
    cols.append(columns.constant_column(A.Uint16, 42, False))
    assert compute_zone_map(eng.lib, eng.space, cols) == want


def test_ranges_of_rle_column_and_synth_batch():
    eng = H.get_backend("b200")
    hb = synth.generate_batch(3, 50_000, num_cities=77, null_rate=0.05)
    import test_pipeline_parity as T
    b = T.upload(eng, hb)
    assert compute_zone_map(eng.lib, eng.space, b.columns) == synth.zone_map(hb)
    # mode 3: values per run
    runs = np.array([5, 9, 9, 200, 7], np.uint16)
    counts = np.array([0, 10, 20, 35, 36, 50], np.uint32)
    buf, vp = columns.make_column(eng.space, A.Uint16, runs, valid=np.array([1, 1, 0, 0, 1], np.uint8), counts=counts)
    assert compute_zone_map(eng.lib, eng.space, [vp]) == {0: (5, 9)}
