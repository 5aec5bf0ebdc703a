"""Archive batches on the host: sorted, run-length encoded columns and the prefilter slicing the reference applies
before a batch is copied to the device.

An archive batch is sorted by its table's archiving sort columns; a sort column is stored as one value per run plus a
cumulative count vector (mode 3), runs of a later sort column nest inside runs of the earlier ones, the remaining
columns hold one value per row.  A query whose filters pin the leading sort columns (equality prefilters, then at most
one range prefilter — query/aql_compiler.go matchPrefilters) does not scan the batch: it binary-searches the matching
row range sort column by sort column and ships only those rows (qc.prefilterSlice, query/aql_processor.go:925-982;
cVectorParty.SliceByValue / SliceIndex, memstore/vector_party.go:371-432).  The sliced batch keeps ABSOLUTE row numbers
in its count vectors; index space = the finest requested column (transferArchiveBatch, query/aql_processor.go:568-626).

This module mirrors that logic on numpy arrays and hands the result to the executors as an ordinary `Batch` (the fused
kernel decodes the RLE columns in place, from their runs), and the scan of a shard's archive batches: which days a time
filter touches, and that only the first and the last of them evaluate it.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import columns
from .executor import Batch

NO_BOUNDARY, INCLUSIVE, EXCLUSIVE = 0, 1, 2   # boundaryType of a range prefilter (query/aql_context.go)


@dataclass
class ArchiveColumn:
    data_type: int
    values: np.ndarray                 # one per run (compressed) or one per row
    valid: np.ndarray | None = None    # same length as values, or None (all valid)
    counts: np.ndarray | None = None   # cumulative row counts, len(values) + 1, counts[0] = first row; None: uncompressed

    @property
    def compressed(self) -> bool:
        return self.counts is not None


def compress(data_type: int, rows: np.ndarray, valid: np.ndarray | None = None) -> ArchiveColumn:
    """Run-length encodes a per-row vector (what archiving does for a sort column)."""
    rows = np.asarray(rows)
    ok = np.ones(len(rows), np.uint8) if valid is None else np.asarray(valid, np.uint8)
    change = np.ones(len(rows), bool)
    change[1:] = (rows[1:] != rows[:-1]) | (ok[1:] != ok[:-1])
    starts = np.flatnonzero(change)
    counts = np.concatenate([starts, [len(rows)]]).astype(np.uint32)
    return ArchiveColumn(data_type, rows[starts].copy(), None if valid is None else ok[starts].copy(), counts)


def slice_index(col: ArchiveColumn, lo_row: int, hi_row: int) -> tuple[int, int]:
    """cVectorParty.SliceIndex: the runs that overlap rows [lo_row, hi_row)."""
    if not col.compressed:
        return lo_row, hi_row
    start = int(np.searchsorted(col.counts, lo_row, side="right")) - 1            # UpperBound(lo) - 1
    end = start + int(np.searchsorted(col.counts[start:], hi_row, side="left"))    # LowerBound(hi) from start
    if end == len(col.counts):
        end -= 1
    return start, end


def slice_by_value(col: ArchiveColumn, lo_row: int, hi_row: int, value) -> tuple[int, int, int, int]:
    """cVectorParty.SliceByValue: (startRow, endRow, startIndex, endIndex) of the rows in [lo_row, hi_row) whose value
    equals `value` (the column is sorted within that row range); an absent value yields the empty range at its
    insertion point, which is what the range prefilter relies on."""
    if col.compressed:
        si, ei = slice_index(col, lo_row, hi_row)
        vals = col.values[si:ei]
        s = si + int(np.searchsorted(vals, value, side="left"))
        e = si + int(np.searchsorted(vals, value, side="right"))
        return int(col.counts[s]), int(col.counts[e]), s, e
    vals = col.values[lo_row:hi_row]
    s = lo_row + int(np.searchsorted(vals, value, side="left"))
    e = lo_row + int(np.searchsorted(vals, value, side="right"))
    return s, e, s, e


@dataclass
class SlicedBatch:
    start_row: int
    end_row: int
    index_ranges: dict          # column -> (startIndex, endIndex)
    first_column: int           # finest requested column: its runs (or rows) are the batch's index space


def prefilter_slice(cols: dict, scan_order: list, num_rows: int, equality_values=(), range_prefilter=None) -> SlicedBatch:
    """qc.prefilterSlice over the requested columns.  `scan_order`: column ids, sort columns first in sort order, then the
    rest; equality_values[k] pins scan_order[k]; range_prefilter = (lower, lower_boundary, upper, upper_boundary) applies
    to the next sort column."""
    start, end = 0, num_rows
    ranges = {}
    for k, c in enumerate(scan_order):
        col = cols[c]
        unmatched = False
        if k < len(equality_values):
            start, end, si, ei = slice_by_value(col, start, end, equality_values[k])
        elif k == len(equality_values) and range_prefilter is not None:
            lower, lower_b, upper, upper_b = range_prefilter
            si, ei = slice_index(col, start, end)
            if lower_b != NO_BOUNDARY:
                ls, le, lsi, lei = slice_by_value(col, start, end, lower)
                start, si = (ls, lsi) if lower_b == INCLUSIVE else (le, lei)
            else:
                unmatched = True
            if upper_b != NO_BOUNDARY:
                us, ue, usi, uei = slice_by_value(col, start, end, upper)
                end, ei = (ue, uei) if upper_b == INCLUSIVE else (us, usi)
            else:
                unmatched = True
        else:
            unmatched = True
        if unmatched:
            si, ei = slice_index(col, start, end)
        ranges[c] = (si, ei)
    return SlicedBatch(start, end, ranges, scan_order[-1])


def to_batch(space, cols: dict, num_columns: int, sl: SlicedBatch) -> Batch:
    """Uploads the sliced vectors (count vectors keep absolute row numbers) and describes them as one Batch."""
    slices, keep = [None] * num_columns, []
    for c, (si, ei) in sl.index_ranges.items():
        col = cols[c]
        valid = None if col.valid is None else col.valid[si:ei]
        counts = col.counts[si:ei + 1] if col.compressed else None
        buf, vp = columns.make_column(space, col.data_type, col.values[si:ei], valid=valid, counts=counts)
        slices[c] = vp
        keep.append(buf)
    for c in range(num_columns):
        if slices[c] is None:                       # a column the query does not read: empty constant slice
            slices[c] = columns.constant_column(cols[c].data_type if c in cols else 0, 0, False)
    first = cols[sl.first_column]
    si, ei = sl.index_ranges[sl.first_column]
    if first.compressed:
        bc = space.put(np.ascontiguousarray(first.counts[si:ei + 1], np.uint32))
        keep.append(bc)
        return Batch(slices, ei - si, base_counts=bc, start_count=0, keep=keep)
    return Batch(slices, sl.end_row - sl.start_row, base_counts=None, start_count=sl.start_row, keep=keep)


# ---- the compiler half: which of a query's filters are prefilters (reference AQLQueryContext.matchPrefilters /
# extractFilter, query/aql_compiler.go:618-765) -----------------------------------------------------------------------
@dataclass
class Prefilters:
    equality_values: list            # one per leading sort column pinned by `column = value`
    range_prefilter: tuple | None    # (lower, lower_boundary, upper, upper_boundary) on the next sort column, or None
    prefilter_ids: list              # indexes into the query's filter list (sorted): these are implied by the slice
    columns: list                    # the sort columns they cover, in sort order


def match_prefilters(filters: list, sort_columns: list) -> Prefilters:
    """Walks the table's archiving sort columns in order: a `column = literal` filter pins the column and matching goes on
    with the next one; otherwise `column > / >= literal` and `column < / <= literal` filters give the (single) range
    prefilter and matching stops; a sort column without a candidate filter stops it too.  Only filters of the form
    `column OP number` with the column on the left qualify, as in the reference; when several filters of one kind
    name the same column the last one wins (the reference's map assignment)."""
    from . import cabi as A
    from . import expr as E
    cand: dict = {}
    for fid, f in enumerate(filters):
        if not isinstance(f, E.Binary) or f.op not in (A.Equal, A.LessThan, A.LessThanOrEqual, A.GreaterThan, A.GreaterThanOrEqual):
            continue
        if not isinstance(f.lhs, E.Col):
            continue
        slot = cand.setdefault(f.lhs.index, [-1, -1, -1])       # lower bound, upper bound, equality
        slot[0 if f.op in (A.GreaterThan, A.GreaterThanOrEqual) else 1 if f.op in (A.LessThan, A.LessThanOrEqual) else 2] = fid

    def extract(fid):
        f = filters[fid]
        if not isinstance(f.rhs, E.Lit):
            return None
        boundary = INCLUSIVE if f.op in (A.GreaterThanOrEqual, A.LessThanOrEqual) else EXCLUSIVE
        return f.rhs.value, boundary

    out = Prefilters([], None, [], [])
    for c in sort_columns:
        if c not in cand:
            break                                               # stop on the first sort column without a filter
        lo, hi, eq = cand[c]
        if eq >= 0:
            v = extract(eq)
            if v is None:
                break
            out.equality_values.append(v[0])
            out.prefilter_ids.append(eq)
            out.columns.append(c)
            continue
        lower = upper = None
        if lo >= 0:
            lower = extract(lo)
        if hi >= 0:
            upper = extract(hi)
        if lower is not None or upper is not None:
            out.range_prefilter = (lower[0] if lower else 0, lower[1] if lower else NO_BOUNDARY,
                                   upper[0] if upper else 0, upper[1] if upper else NO_BOUNDARY)
            out.prefilter_ids += [i for i, x in ((lo, lower), (hi, upper)) if x is not None]
            out.columns.append(c)
        break                                                   # stop after the first range filter
    out.prefilter_ids.sort()
    return out


# ---- which archive batches a query scans, and which of them evaluate its time filter -------------------------------
SECONDS_PER_DAY = 86400


def archive_batch_ids(time_range, now: int) -> range:
    """Archive batch IDs (= days since the epoch) a query with time filter [from, to) scans: from / 86400 up to
    (to + 86399) / 86400, open ends = day 0 / the day of `now` (TableScanner.ArchiveBatchIDStart / End, reference
    query/aql_compiler.go:1044-1057)."""
    frm, to = time_range
    start = 0 if frm is None else int(frm) // SECONDS_PER_DAY
    end = (int(now if to is None else to) + SECONDS_PER_DAY - 1) // SECONDS_PER_DAY
    return range(start, max(end, start))


def scan_archive_batches(executor, batches: dict, time_range, now: int) -> list:
    """Feeds the resident archive batches {day: Batch} of a shard to `executor` the way processShard does
    (query/aql_processor.go:222-248): days in the scanned range, empty ones skipped, and the time filter evaluated for the
    FIRST and the LAST scanned day only — every row of a day in between lies inside [from, to) by construction (an archive
    batch holds the rows of its day, and the fact table's time column is never NULL).
    Returns [(day, evaluated the time filter)] of the batches processed."""
    ids = archive_batch_ids(time_range, now)
    done = []
    for day in ids:
        b = batches.get(day)
        if b is None or b.num_rows == 0:
            continue
        first_or_last = day == ids.start or day == ids.stop - 1
        executor.process_batch(b, time_filters=first_or_last)
        done.append((day, first_or_last))
    return done


def scan_shard(executor, live_batches: list, archive_batches: dict, cutoff: int, time_range, now: int) -> dict:
    """processShard (query/aql_processor.go:166-248): the live batches first — scanned when the range reaches past the
    archiving cutoff, each with the cutoff filter `time >= cutoff` (rows below it are the archive's) —, then the archive days
    when the range starts below the cutoff.  Zone-map skipping of live batches happens inside the executor."""
    frm, to = time_range
    done = {"live": 0, "archive": []}
    if to is None or cutoff < to:
        for b in live_batches:
            if b.num_rows:
                executor.process_batch(b, cutoff=cutoff)
                done["live"] += 1
    if archive_batches and (frm is None or cutoff > frm):
        done["archive"] = scan_archive_batches(executor, archive_batches, time_range, now)
    return done
