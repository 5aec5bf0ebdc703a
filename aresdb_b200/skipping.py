"""Batch skipping by zone map: a batch whose per-column min / max (executor.Batch.ranges, the same numbers that go to
the engine as BatchPlan.Ranges) contradict a `column OP literal` filter of the query cannot contribute a row and is not
sent to the GPU at all.

This is the reference's live-batch skipping (shouldSkipLiveBatch / shouldSkipLiveBatchWithFilter,
query/aql_processor.go:1435-1526): filter on the main table, binary, OP one of EQ / GTE / GT / LTE / LT, one side a
column reference and the other an integer literal (either order — the operator is mirrored when the literal is on the
left).  The reference restricts it to the Uint32 time column because that is the only column it keeps min / max for
(LiveVectorParty.GetMinMaxValue); here every integer column with a zone-map entry qualifies.  NULL rows never pass a
comparison, so only the valid values' range matters.
"""
from __future__ import annotations

from . import cabi as A
from . import expr as E

_MIRROR = {A.GreaterThanOrEqual: A.LessThanOrEqual, A.GreaterThan: A.LessThan,
           A.LessThanOrEqual: A.GreaterThanOrEqual, A.LessThan: A.GreaterThan, A.Equal: A.Equal}


def filter_excludes_range(f: E.Expr, ranges: dict) -> bool:
    """True when no valid value inside the zone map can satisfy filter `f`."""
    if not isinstance(f, E.Binary) or f.op not in _MIRROR:
        return False
    op, col, lit = f.op, f.lhs, f.rhs
    if isinstance(col, E.Lit) and isinstance(lit, E.Col):       # literal on the left: swap and mirror the operator
        col, lit, op = lit, col, _MIRROR[op]
    if not (isinstance(col, E.Col) and isinstance(lit, E.Lit)) or lit.type == E.Type.Float:
        return False
    if col.type not in (E.Type.Unsigned, E.Type.Signed, E.Type.Boolean) or col.index not in ranges:
        return False
    lo, hi = ranges[col.index]
    num = int(lit.value)
    # The kernel (and the reference) compares in the promoted class: an integer literal is a ConstInt (int32), so the
    # comparison runs in int32 whatever the column is (query/utils.hpp:83-94) and a uint32 value >= 2^31 compares as a
    # negative number.  Only skip when the literal and the whole range are representable in int32 without wrapping, so
    # that the unbounded comparison below says what the kernel would.
    if not (-(2 ** 31) <= num < 2 ** 31 and 0 <= lo <= hi < 2 ** 31):
        return False
    if op == A.GreaterThanOrEqual:
        return hi < num
    if op == A.GreaterThan:
        return hi <= num
    if op == A.LessThanOrEqual:
        return lo > num
    if op == A.LessThan:
        return lo >= num
    return lo > num or hi < num                                  # Equal


def should_skip_batch(query, ranges: dict | None) -> bool:
    """`query.filters` are ANDed: one filter the zone map contradicts is enough."""
    if not ranges:
        return False
    return any(filter_excludes_range(f, ranges) for f in query.filters)
