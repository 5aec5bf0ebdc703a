"""The `joins` clause of the AQL front-end (aresdb_b200/aql.py process_joins): acceptance rules of the reference's
processJoinConditions / matchEqualJoin (query/aql_compiler.go:168-282; its cases: query/aql_compiler_test.go:1552-1770),
`alias.column` operands, and — end to end on the CPU backends — the compiled query against the hand-built one of
tests/test_joins.py on the reference's HOST build and the oracle."""
import pytest

import harness as H
import test_joins as TJ
import test_pipeline_parity as T
from aresdb_b200 import aql, cabi as A, expr as E, synth


def _schemas():
    trips = aql.Table("trips", [aql.Column("request_at", A.Uint32), aql.Column("city_id", A.Uint16), aql.Column("status", A.Uint8),
                                aql.Column("fare", A.Float32)])
    cities = aql.Table("api_cities", [aql.Column("id", A.Uint16), aql.Column("region", A.Uint8, enum={"emea": 2, "apac": 5}),
                                      aql.Column("tz", A.Uint8), aql.Column("surge", A.Float32)],
                       primary_key=["id"], is_fact_table=False)
    return trips, cities


def _query(joins, dims=None, filters=None, measure="count(*)"):
    return {"table": "trips", "joins": joins, "measures": [{"sqlExpression": measure}], "rowFilters": filters or [],
            "dimensions": dims or [{"sqlExpression": "city_id"}]}


def test_equal_join_is_accepted_in_either_order_and_under_an_alias():
    trips, cities = _schemas()
    known = {"api_cities": aql.JoinedTable(cities, resident="resident table")}
    for cond, alias in (("city_id = api_cities.id", None), ("api_cities.id = trips.city_id", None), ("c.id = city_id", "c")):
        j = {"table": "api_cities", "conditions": [cond]}
        if alias:
            j["alias"] = alias
        prefix = alias or "api_cities"
        q = aql.compile_query(_query([j], dims=[{"sqlExpression": f"{prefix}.region"}, {"sqlExpression": "status"}],
                                     filters=[f"{prefix}.surge > 1.5", "status = 1", f"{prefix}.region = 'apac'"]), trips, 0,
                              dimension_tables=known)
        assert len(q.joins) == 1 and q.joins[0].table == "resident table"
        assert isinstance(q.joins[0].on, E.Col) and q.joins[0].on.index == 1          # trips.city_id
        region = q.dimensions[0]
        assert isinstance(region, E.ForeignCol) and (region.table, region.index, region.data_type) == (0, 1, A.Uint8)
        # main-table filters first, filters that read the joined table after the join (aql_batchexecutor.go:100-147)
        assert q.num_main_filters == 1 and not E.uses_foreign(q.filters[0]) and all(E.uses_foreign(f) for f in q.filters[1:])
        enum_filter = q.filters[2]
        assert isinstance(enum_filter.rhs, E.Lit) and enum_filter.rhs.value == 5       # 'apac' through the JOINED column's dictionary


def test_join_conditions_the_reference_rejects():
    trips, cities = _schemas()
    fact = aql.Table("other_trips", [aql.Column("id", A.Uint16)], primary_key=["id"], is_fact_table=True)
    composite = aql.Table("pairs", [aql.Column("a", A.Uint16), aql.Column("b", A.Uint16)], primary_key=["a", "b"], is_fact_table=False)
    regions = aql.Table("regions", [aql.Column("id", A.Uint8), aql.Column("name", A.Uint8)], primary_key=["id"], is_fact_table=False)
    known = {t.name: aql.JoinedTable(t) for t in (cities, fact, composite, regions)}
    ok = {"table": "api_cities", "conditions": ["city_id = api_cities.id"]}

    def bad(joins, match):
        with pytest.raises(aql.AQLError, match=match):
            aql.compile_query(_query(joins), trips, 0, dimension_tables=known)

    bad([{"table": "api_cities", "conditions": ["city_id = api_cities.id", "status = api_cities.tz"]}], "1 join conditions expected, got 2")
    bad([{"table": "api_cities", "conditions": []}], "1 join conditions expected, got 0")
    bad([{"table": "api_cities", "conditions": ["city_id > api_cities.id"]}], "equal join expected")
    bad([{"table": "api_cities", "conditions": ["city_id + 1 = api_cities.id"]}], "column in join condition expected")
    bad([{"table": "api_cities", "conditions": ["api_cities.id"]}], "binary expression expected")
    bad([{"table": "other_trips", "conditions": ["city_id = other_trips.id"]}], "fact table")
    bad([{"table": "pairs", "conditions": ["city_id = pairs.a"]}], "composite key")
    bad([{"table": "api_cities", "conditions": ["status = api_cities.tz"]}], "not primary key")
    bad([{"table": "api_cities", "conditions": ["city_id = city_id"]}], "joined directly to the main table")
    # a table joined THROUGH another joined table instead of the main one
    bad([ok, {"table": "regions", "conditions": ["api_cities.region = regions.id"]}], "joined directly to the main table")
    bad([{"table": "nowhere", "conditions": ["city_id = nowhere.id"]}], "unknown table")
    bad([{"table": "api_cities", "conditions": ["geography_intersects(city_id, api_cities.id)"]}], "geo joins")
    bad([dict(ok, alias=f"c{i}", conditions=[f"city_id = c{i}.id"]) for i in range(9)], "At most 8 foreign tables allowed, got: 9")
    # eight are fine
    q = aql.compile_query(_query([dict(ok, alias=f"c{i}", conditions=[f"city_id = c{i}.id"]) for i in range(8)],
                                 dims=[{"sqlExpression": "c7.region"}]), trips, 0, dimension_tables=known)
    assert len(q.joins) == 8 and q.dimensions[0].table == 7
    with pytest.raises(aql.AQLError, match="unknown table"):
        aql.compile_query(_query([ok], dims=[{"sqlExpression": "cities.region"}]), trips, 0, dimension_tables=known)


@pytest.mark.parametrize("backend", ["ref", "oracle"])
def test_compiled_join_query_equals_the_hand_built_one(backend):
    """AQL text -> AggQuery with a join == the query tests/test_joins.py builds by hand, through the legacy call sequence
    (HashLookup + ForeignColumnInput) on the reference's HOST build and on the oracle."""
    be = H.get_backend(backend)
    table, _ = TJ._dimension_table(be)
    tzb, tzn = TJ._tz_table(be)
    trips, cities = _schemas()
    known = {"api_cities": aql.JoinedTable(cities, resident=table)}
    hbs = [synth.generate_batch(d, n, num_cities=80, null_rate=0.03) for d, n in ((0, 6000), (1, 2500))]
    text = {"table": "trips", "joins": [{"table": "api_cities", "alias": "c", "conditions": ["trips.city_id = c.id"]}],
            "measures": [{"sqlExpression": "sum(fare)", "rowFilters": ["status = 1"]}], "rowFilters": ["c.surge > 1.0"],
            "dimensions": [{"sqlExpression": "c.region"}, {"sqlExpression": "request_at", "timeBucketizer": "hour"}]}
    compiled = aql.compile_query(text, trips, 0, dimension_tables=known)
    by_hand = TJ.join_queries(table, tzb.ptr, tzn)["by_region"]
    got, exp = T.run_legacy(be, compiled, hbs), T.run_legacy(be, by_hand, hbs)
    assert exp.groups > 0
    T.assert_same_result(got, exp, ctx="by_region")
    # a measure and an IS NULL filter over joined columns
    text2 = {"table": "trips", "joins": [{"table": "api_cities", "conditions": ["city_id = api_cities.id"]}],
             "measures": [{"sqlExpression": "sum(api_cities.surge)"}], "dimensions": [{"sqlExpression": "status"}]}
    got2 = T.run_legacy(be, aql.compile_query(text2, trips, 0, dimension_tables=known), hbs)
    T.assert_same_result(got2, T.run_legacy(be, TJ.join_queries(table, tzb.ptr, tzn)["sum_surge"], hbs), ctx="sum_surge")


ZONES = ["America/Los_Angeles", "America/New_York", "Europe/Amsterdam", "Asia/Kolkata", "Asia/Tokyo", "Australia/Sydney",
         "America/Sao_Paulo", "Africa/Johannesburg", "Asia/Kathmandu", "Pacific/Auckland", "UTC", "America/St_Johns"]


@pytest.mark.parametrize("backend", ["ref", "oracle"])
def test_time_zone_column_joins_the_timezone_table_and_shifts_by_the_joined_offset(backend):
    """`"timezone": "tz(city_id)"`: the configured timezone table joins on city_id = __timezone_lookup.id, the enum column
    `tz` (dictionary of IANA names) becomes an int16 offset table at `now`, and the hour bucket is taken in the joined city's
    local time — the hand-built `local_hour` query of tests/test_joins.py with the same offset table."""
    import datetime as dt
    import zoneinfo
    import numpy as np
    be = H.get_backend(backend)
    table, _ = TJ._dimension_table(be)
    trips, cities = _schemas()
    cities.columns[2] = aql.Column("tz", A.Uint8, enum={z: i for i, z in enumerate(ZONES)})
    known = {"api_cities": aql.JoinedTable(cities, resident=table)}
    now = 1_720_000_000          # July: daylight saving in the northern zones
    offsets = np.array([int(dt.datetime.fromtimestamp(now, zoneinfo.ZoneInfo(z)).utcoffset().total_seconds()) for z in ZONES],
                       np.int64).astype(np.int16)   # int16(offset) as in the reference: +10 h and +12 h wrap
    assert offsets[0] == -7 * 3600 and offsets[3] == 19800 and offsets[8] == 20700 and offsets[11] == -9000
    assert offsets[5] == 36000 - 65536 and offsets[9] == 43200 - 65536
    hbs = [synth.generate_batch(d, n, num_cities=80, null_rate=0.03) for d, n in ((0, 6000), (1, 2500))]
    text = {"table": "trips", "timezone": "tz(city_id)", "measures": [{"sqlExpression": "count(*)"}], "rowFilters": ["city_id != 0"],
            "dimensions": [{"sqlExpression": "request_at", "timeBucketizer": "hour"}]}
    uploads = []

    def upload(a):
        uploads.append(a.copy())
        return be.put(a)

    q = aql.compile_query(text, trips, now, dimension_tables=known, timezone_table="api_cities", upload=upload)
    assert len(q.joins) == 1 and q.joins[0].on.index == 1 and q.joins[0].timezone_size == 12 and (uploads[0] == offsets).all()
    tzb = be.put(offsets)
    by_hand = TJ.join_queries(table, tzb.ptr, len(offsets))["local_hour"]
    got, exp = T.run_legacy(be, q, hbs), T.run_legacy(be, by_hand, hbs)
    assert exp.groups > 24
    T.assert_same_result(got, exp, ctx="local_hour")
    # the query joins the timezone table itself: its alias is used, no second join
    text2 = dict(text, joins=[{"table": "api_cities", "alias": "c", "conditions": ["city_id = c.id"]}])
    q2 = aql.compile_query(text2, trips, now, dimension_tables=known, timezone_table="api_cities", upload=upload)
    assert len(q2.joins) == 1
    T.assert_same_result(T.run_legacy(be, q2, hbs), exp, ctx="local_hour, explicit join")
    for bad, match in ((dict(text, timezone="nope(city_id)"), "unknown timezone column"),
                       (dict(text, timezone="region(city_id)"), "error parsing timezone"),
                       (dict(text, timezone="surge(city_id)"), "unknown timezone column")):
        with pytest.raises(aql.AQLError, match=match):
            aql.compile_query(bad, trips, now, dimension_tables=known, timezone_table="api_cities", upload=upload)
    with pytest.raises(aql.AQLError, match="configured timezone table"):
        aql.compile_query(text, trips, now, dimension_tables=known, upload=upload)


@pytest.mark.gpu
def test_compiled_join_and_time_zone_column_queries_on_the_fused_path():
    """The same AQL texts through ExecuteBatchPlan on the GPU (the lookup is a gather stage of the fused kernel, the
    time-zone offset table a plan operand) against the reference call sequence on the oracle."""
    import datetime as dt
    import zoneinfo
    import numpy as np
    eng, orc = H.get_backend("b200"), H.get_backend("oracle")
    trips, cities = _schemas()
    cities.columns[2] = aql.Column("tz", A.Uint8, enum={z: i for i, z in enumerate(ZONES)})
    now = 1_720_000_000
    hbs = [synth.generate_batch(d, n, num_cities=80, null_rate=0.03) for d, n in ((0, 20000), (1, 7777))]
    texts = {
        "by_region": {"table": "trips", "joins": [{"table": "api_cities", "alias": "c", "conditions": ["trips.city_id = c.id"]}],
                      "measures": [{"sqlExpression": "sum(fare)", "rowFilters": ["status = 1"]}], "rowFilters": ["c.surge > 1.0", "c.region in (1, 2, 5)"],
                      "dimensions": [{"sqlExpression": "c.region"}, {"sqlExpression": "request_at", "timeBucketizer": "hour"}]},
        "local_hour": {"table": "trips", "timezone": "tz(city_id)", "measures": [{"sqlExpression": "count(*)"}], "rowFilters": ["city_id != 0"],
                       "dimensions": [{"sqlExpression": "request_at", "timeBucketizer": "hour"}]},
    }
    res = {}
    for be in (orc, eng):
        table, _ = TJ._dimension_table(be)
        known = {"api_cities": aql.JoinedTable(cities, resident=table)}
        for name, text in texts.items():
            q = aql.compile_query(text, trips, now, dimension_tables=known, timezone_table="api_cities", upload=be.put)
            res[(name, be is eng)] = T.run_fused(be, q, hbs) if be is eng else T.run_legacy(be, q, hbs)
    for name in texts:
        assert res[(name, False)].groups > 10
        T.assert_same_result(res[(name, True)], res[(name, False)], ctx=name)
