"""Dimension-table joins, host side.

The reference joins a fact table with small dimension tables through the memstore's primary-key index of the
dimension table: per batch, HashLookup turns the main table's join column into one RecordID (batch id, row) per index
position (query/aql_batchexecutor.go:115-147), and every later filter / dimension / measure may read dimension-table
columns at those RecordIDs (ForeignColumnInput, query/time_series_aggregate.go:96-125).  This module builds what the
memstore hands to that path: the column batches of a dimension table in device memory and its cuckoo hash index in the
byte layout HashLookupFunctor probes (memstore/cuckoo_index.go:42-48: per bucket RecordID[8] | signature[8] | key[8],
`numBuckets` buckets followed by one stash bucket of which 4 cells are used)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import cabi as A
from . import columns

BUCKET_CELLS, STASH_CELLS = 8, 4
BASE_BATCH_ID = -2147483648   # memstore.BaseBatchID (live batch ids start at INT32_MIN; 0 means "no record")
DEFAULT_SEEDS = (2596996162, 4039455774, 2854263694, 1879968118)


def murmur3_32(key: bytes, seed: int) -> int:
    """MurmurHash3 x86_32 (the reference's murmur3sum32, query/utils.cu:113-155)."""
    M = 0xFFFFFFFF
    h, n = seed & M, len(key)
    for i in range(0, n - n % 4, 4):
        k = int.from_bytes(key[i:i + 4], "little")
        k = (k * 0xcc9e2d51) & M
        k = ((k << 15) | (k >> 17)) & M
        k = (k * 0x1b873593) & M
        h ^= k
        h = ((h << 13) | (h >> 19)) & M
        h = (h * 5 + 0xe6546b64) & M
    tail = key[n - n % 4:]
    if tail:
        k = int.from_bytes(tail, "little")
        k = (k * 0xcc9e2d51) & M
        k = ((k << 15) | (k >> 17)) & M
        k = (k * 0x1b873593) & M
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85ebca6b) & M
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & M
    return h ^ (h >> 16)


def build_cuckoo_index(keys: list, record_ids: list, seeds=DEFAULT_SEEDS, num_hashes: int = 4):
    """Places every (key bytes, (batch id, row)) in the first free cell of its candidate buckets, else in the stash,
    growing the bucket count until everything fits — any placement HashLookupFunctor finds is a valid index (the
    memstore's own insert evicts and rehashes; lookups do not care how a key got to its cell).
    Returns (uint8 array, seeds, key_bytes, num_hashes, num_buckets)."""
    assert len(keys) == len(record_ids) and len(keys) > 0
    kb = len(keys[0])
    assert all(len(k) == kb for k in keys) and len(set(keys)) == len(keys)
    cell = 8 + kb + 1
    bucket_bytes = BUCKET_CELLS * cell
    off_sig, off_key = BUCKET_CELLS * 8, BUCKET_CELLS * 8 + BUCKET_CELLS
    num_buckets = max(2, (len(keys) + BUCKET_CELLS - 1) // BUCKET_CELLS * 2)
    while True:
        raw = np.zeros(bucket_bytes * (num_buckets + 1), np.uint8)
        used = np.zeros((num_buckets + 1, BUCKET_CELLS), bool)
        ok = True
        for key, (batch_id, row) in zip(keys, record_ids):
            placed = False
            for t in range(num_hashes):
                h = murmur3_32(key, seeds[t])
                b = h % num_buckets
                free = np.nonzero(~used[b])[0]
                if free.size:
                    j, sig, base = int(free[0]), max(h >> 24, 1), b * bucket_bytes
                    placed = True
                    break
            if not placed:
                free = np.nonzero(~used[num_buckets, :STASH_CELLS])[0]
                if not free.size:
                    ok = False
                    break
                j, sig, b, base = int(free[0]), max(murmur3_32(key, seeds[0]) >> 24, 1), num_buckets, num_buckets * bucket_bytes
            used[b, j] = True
            raw[base + 8 * j: base + 8 * j + 8] = np.frombuffer(np.array([batch_id], "<i4").tobytes() + np.array([row], "<u4").tobytes(), np.uint8)
            raw[base + off_sig + j] = sig
            raw[base + off_key + j * kb: base + off_key + (j + 1) * kb] = np.frombuffer(key, np.uint8)
        if ok:
            return raw, tuple(seeds), kb, num_hashes, num_buckets
        num_buckets = num_buckets * 3 // 2 + 1


@dataclass
class DimensionTable:
    """One dimension table resident in a memory space: per column the batches' slices + the primary-key index."""
    column_types: list
    batches: list            # list of batches; a batch = list of VectorPartySlice (one per column)
    num_records_in_last_batch: int
    index_buf: object        # Buf holding the cuckoo index
    seeds: tuple
    key_bytes: int
    num_hashes: int
    num_buckets: int
    defaults: list           # per column cabi.DefaultValue
    base_batch_id: int = BASE_BATCH_ID
    keep: list = field(default_factory=list)

    @classmethod
    def build(cls, space, column_types: list, values: list, valid: list | None, pk_column: int, rows_per_batch: int | None = None,
              base_batch_id: int = BASE_BATCH_ID):
        """values / valid: one array per column over all rows; rows are cut into batches of rows_per_batch."""
        n = len(values[0])
        rows_per_batch = rows_per_batch or n
        batches, keep = [], []
        for b0 in range(0, n, rows_per_batch):
            cols = []
            for c, dt in enumerate(column_types):
                ok = None if valid is None or valid[c] is None else np.asarray(valid[c])[b0:b0 + rows_per_batch]
                buf, vp = columns.make_column(space, dt, np.asarray(values[c])[b0:b0 + rows_per_batch], valid=ok)
                keep.append(buf)
                cols.append(vp)
            batches.append(cols)
        pk = np.ascontiguousarray(values[pk_column])
        kb = pk.dtype.itemsize if pk.ndim == 1 else pk.shape[1] * pk.dtype.itemsize
        keys = [pk[i].tobytes() for i in range(n)]
        # batch ids must be non-zero for a record to count as found (RecordID {0, 0} = not found): memstore live batch
        # ids start at BaseBatchID = INT32_MIN and grow
        rids = [(base_batch_id + i // rows_per_batch, i % rows_per_batch) for i in range(n)]
        assert all(r[0] != 0 for r in rids)
        raw, seeds, kb2, nh, nb = build_cuckoo_index(keys, rids)
        assert kb2 == kb
        ibuf = space.put(raw)
        last = n - (len(batches) - 1) * rows_per_batch
        defaults = [A.make_default_value(False, 0, dt) for dt in column_types]
        return cls(list(column_types), batches, last, ibuf, seeds, kb, nh, nb, defaults, base_batch_id, keep)

    def hash_index(self) -> A.CuckooHashIndex:
        h = A.CuckooHashIndex()
        h.buckets = self.index_buf.ptr
        for i in range(4):
            h.seeds[i] = self.seeds[i]
        h.keyBytes, h.numHashes, h.numBuckets = self.key_bytes, self.num_hashes, self.num_buckets
        return h

    def column_slices(self, column: int):
        """ctypes array of this column's VectorPartySlice, one per batch (kept alive by the caller)."""
        arr = (A.VectorPartySlice * len(self.batches))(*[b[column] for b in self.batches])
        return arr

    def foreign_column(self, column: int, record_ids_ptr: int | None, tz_ptr: int | None = None, tz_size: int = 0):
        """(ForeignColumnVector, keep-alive) of one column read at `record_ids_ptr`."""
        arr = self.column_slices(column)
        f = A.ForeignColumnVector()
        f.RecordIDs = record_ids_ptr
        f.Batches = C.cast(arr, C.c_void_p).value
        f.BaseBatchID, f.NumBatches = self.base_batch_id, len(self.batches)
        f.NumRecordsInLastBatch = self.num_records_in_last_batch
        f.TimezoneLookup, f.TimezoneLookupSize = tz_ptr, tz_size
        f.DataType = self.column_types[column]
        f.DefaultValue = self.defaults[column]
        return f, arr


def foreign_input(table: DimensionTable, column: int, record_ids_ptr: int, tz_ptr: int | None = None, tz_size: int = 0):
    """InputVector of a dimension-table column (makeForeignColumnInput, query/time_series_aggregate.go:96-125)."""
    f, keep = table.foreign_column(column, record_ids_ptr, tz_ptr, tz_size)
    iv = A.InputVector()
    iv.Vector.ForeignVP = f
    iv.Type = A.ForeignColumnInput
    return iv, keep
