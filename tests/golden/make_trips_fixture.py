#!/usr/bin/env python
"""Builds tests/golden/trips_1k.npz from the reference's example data set (BASELINE config 1:
examples/1k_trips, count / sum group-by hour on QUERY_MODE=HOST).

Runs ONLY in the development container (needs /root/reference and oracle/_ref); the GPU box and the
test-suite read the committed .npz.  What is stored:
  * the columns of examples/1k_trips/data/trips.csv the two example queries touch (city_id, status as
    dictionary ids in order of first appearance, fare) and a request_at column generated the way
    examples/utils/utils.go:40-55 materialises the "{1d}" placeholder (now - 1 day + uniform offset), with a
    fixed seed and a fixed `now` so that the fixture is reproducible;
  * for each example query (total_trips.aql: count(*), total_fare.aql: sum(fare); both
    status='completed', last 24 hours, bucketed by hour) the result of the REFERENCE's HOST build of the
    hot path, driven through its C ABI with the plan aresdb_b200.aql compiles from the query.
"""
import csv
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

NOW = 1_700_012_345          # 2023-11-15 01:39:05 UTC
SEED = 20260922


def queries():
    """The two example queries (same content as examples/1k_trips/queries/total_{trips,fare}.aql)."""
    def q(measure):
        return {"table": "trips",
                "measures": [{"alias": "value", "sqlExpression": measure, "rowFilters": ["status='completed'"]}],
                "timeFilter": {"column": "request_at", "from": "24 hours ago", "to": "this quarter-hour"},
                "dimensions": [{"alias": "ts", "sqlExpression": "request_at", "timeBucketizer": "hour"}],
                "joins": []}
    return {"total_trips": q("count(*)"), "total_fare": q("sum(fare)")}


def trips_table(status_names):
    from aresdb_b200 import aql, cabi as A
    return aql.Table("trips", [aql.Column("request_at", A.Uint32), aql.Column("city_id", A.Uint16),
                               aql.Column("status", A.Uint8, enum={n: i for i, n in enumerate(status_names)}),
                               aql.Column("fare", A.Float32)])


def main():
    import harness as H
    from aresdb_b200 import aql, columns
    from aresdb_b200.executor import Batch, LegacyBatchExecutor
    src = Path("/root/reference/examples/1k_trips/data/trips.csv")
    rows = list(csv.DictReader(src.open()))
    status_names = []
    for r in rows:
        if r["status"] not in status_names:
            status_names.append(r["status"])
    rng = np.random.default_rng(SEED)
    cols = {
        "request_at": (NOW - 86400 + rng.integers(0, 86400, len(rows))).astype(np.uint32),
        "city_id": np.array([int(r["city_id"]) for r in rows], np.uint16),
        "status": np.array([status_names.index(r["status"]) for r in rows], np.uint8),
        "fare": np.array([float(r["fare"]) for r in rows], np.float32),
    }
    table = trips_table(status_names)
    ref = H.get_backend("ref")
    out = dict(cols, status_names=np.array(status_names), now=np.int64(NOW))
    for name, q in queries().items():
        aq = aql.compile_query(q, table, NOW)
        vps, keep = [], []
        for c in table.columns:
            buf, vp = columns.make_column(ref.space, c.data_type, cols[c.name])
            vps.append(vp)
            keep.append(buf)
        ex = LegacyBatchExecutor(ref.lib, ref.space, aq)
        ex.process_batch(Batch(vps, len(rows)), is_last=True)
        res = ex.result()
        out[f"{name}_hours"] = np.array(res.decoded_dims()[0], np.uint32)
        out[f"{name}_values"] = res.measures
        print(name, res.groups, "groups", res.measures[:5])
    np.savez_compressed(ROOT / "tests" / "golden" / "trips_1k.npz", **out)


if __name__ == "__main__":
    main()
