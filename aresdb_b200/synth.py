"""Synthetic fact table of the benchmark configs (mirrors the schema of the
reference's examples/1k_trips/schema/trips.json):

    request_at u32   base + U[0, 86400) per day-batch (one archive-like batch per UTC day, unsorted)
    city_id    u16   uniform over 1..num_cities (0 never occurs; `city_id != 0` keeps every row)
    status     u8    small enum, 4 values, P(completed = 1) = 0.5
    fare       f32   multiples of 1/64 in [0, 100): double sums are exact in ANY association order,
                     so SUM(fare) is bit-comparable between the GPU, the oracle and the reference

All columns are mode 2 (null bitmap present, `null_rate` nulls, value 0 stored under a NULL).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import cabi as A

COL_REQUEST_AT, COL_CITY_ID, COL_STATUS, COL_FARE = 0, 1, 2, 3
COLUMN_TYPES = [A.Uint32, A.Uint16, A.Uint8, A.Float32]
COLUMN_NAMES = ["request_at", "city_id", "status", "fare"]
BASE_TS = 1_726_963_200  # 2024-09-22T00:00:00Z, a day boundary
STATUS_COMPLETED = 1


@dataclass
class HostBatch:
    """One batch in host memory: per column (values ndarray, valid ndarray[u8])."""
    values: list
    valid: list
    num_rows: int
    day: int


def zipf_cdf(num_cities: int, s: float = 1.1) -> np.ndarray:
    """CDF of P(city = k) ~ k^-s over 1..num_cities (SURVEY.md 8d names a Zipf city distribution: a few hot
    groups take most rows, which is what stresses same-address atomics)."""
    w = np.arange(1, num_cities + 1, dtype=np.float64) ** -s
    return np.cumsum(w / w.sum())


def generate_batch(day: int, rows: int, num_cities: int = 100, null_rate: float = 0.01, seed: int = 20260922,
                   exact_fares: bool = True, city_dist: str = "uniform") -> HostBatch:
    rng = np.random.default_rng([seed, day])
    ts = (BASE_TS + day * 86400 + rng.integers(0, 86400, rows, dtype=np.uint32)).astype(np.uint32)
    city = rng.integers(1, num_cities + 1, rows, dtype=np.uint16).astype(np.uint16)
    if city_dist == "zipf":
        city = (np.searchsorted(zipf_cdf(num_cities), rng.random(rows), side="right") + 1).clip(1, num_cities).astype(np.uint16)
    status = (rng.integers(0, 2, rows, dtype=np.uint8) * rng.integers(1, 4, rows, dtype=np.uint8)).astype(np.uint8)
    # P(status == 1): half the rows get 0 ("not completed" bucket), the rest split 1..3 -> make 1 dominant
    status = np.where(rng.random(rows) < 0.5, np.uint8(STATUS_COMPLETED), status).astype(np.uint8)
    if exact_fares:
        fare = (rng.integers(0, 6400, rows, dtype=np.int32) / 64.0).astype(np.float32)
    else:
        fare = (rng.random(rows, dtype=np.float32) * np.float32(100.0)).astype(np.float32)
    values = [ts, city, status, fare]
    valid = []
    for i, v in enumerate(values):
        ok = (rng.random(rows, dtype=np.float32) >= null_rate).astype(np.uint8)
        values[i] = np.where(ok != 0, v, v.dtype.type(0)).astype(v.dtype)
        valid.append(ok)
    return HostBatch(values, valid, rows, day)


def zone_map(hb: HostBatch) -> dict:
    """{column index: (min, max)} over the VALID values of the integer columns of a batch — what the
    memstore keeps per batch (LiveVectorParty.GetMinMaxValue for the time column; archive day, enum
    dictionary size) and hands to the engine as BatchPlan.Ranges."""
    out = {}
    for i, (v, ok) in enumerate(zip(hb.values, hb.valid)):
        sel = v[ok != 0]
        if not sel.size:
            continue
        if v.dtype == np.float32:   # non-negative finite floats: the IEEE bit patterns (same order as the values)
            if np.isfinite(sel).all() and not np.signbit(sel).any():
                out[i] = (int(sel.min().view(np.uint32)), int(sel.max().view(np.uint32)))
        elif v.dtype.kind in "ui" and int(sel.min()) >= 0 and int(sel.max()) < 2 ** 31:
            out[i] = (int(sel.min()), int(sel.max()))
    return out


def zone_map_of_day(day: int, num_cities: int = 100) -> dict:
    """Zone map of a generated day-batch by construction (generate_batch / generate_batch_cuda): what the
    archive store knows without looking at the data — batch ID = day, city ids 1..num_cities, 4 status values."""
    return {COL_REQUEST_AT: (BASE_TS + day * 86400, BASE_TS + day * 86400 + 86399), COL_CITY_ID: (1, num_cities),
            COL_STATUS: (0, 3), COL_FARE: (0, int(np.float32(100.0).view(np.uint32)))}   # fares in [0, 100): float bits


# ---- large-scale generation on the GPU (bench.py): same schema, torch RNG -------------------------
def generate_batch_cuda(day: int, rows: int, device, num_cities: int = 100, null_rate: float = 0.01,
                        seed: int = 20260922, exact_fares: bool = True, city_dist: str = "uniform"):
    """Returns a list of per-column byte tensors on `device`, each laid out [null bitmap][values]
    with 64-byte aligned parts (mode 2), plus (nulls_offset=0, values_offset) per column.
    Generation is chunked so that temporaries stay small."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000 + day)
    chunk = 1 << 24
    np_bytes = [4, 2, 1, 4]
    bitmap_bytes = (rows + 7) // 8 + 1
    values_off = (bitmap_bytes + 63) // 64 * 64
    bufs = [torch.zeros(values_off + ((rows * w + 64) // 64 * 64), dtype=torch.uint8, device=device) for w in np_bytes]
    weights = (2 ** torch.arange(8, device=device, dtype=torch.int32)).to(torch.uint8)
    cdf = torch.from_numpy(zipf_cdf(num_cities)).to(device) if city_dist == "zipf" else None
    for c0 in range(0, rows, chunk):
        n = min(chunk, rows - c0)
        ts = (torch.randint(0, 86400, (n,), generator=g, device=device, dtype=torch.int64) + (BASE_TS + day * 86400))
        city = torch.randint(1, num_cities + 1, (n,), generator=g, device=device, dtype=torch.int32)
        if cdf is not None:
            u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
            city = (torch.searchsorted(cdf, u, right=True) + 1).clamp(1, num_cities).to(torch.int32)
        st_other = torch.randint(0, 4, (n,), generator=g, device=device, dtype=torch.int32)
        status = torch.where(torch.rand(n, generator=g, device=device) < 0.5,
                             torch.full_like(st_other, STATUS_COMPLETED), st_other)
        if exact_fares:
            fare = torch.randint(0, 6400, (n,), generator=g, device=device, dtype=torch.int32).to(torch.float32) / 64.0
        else:
            fare = torch.rand(n, generator=g, device=device) * 100.0
        cols = [ts.to(torch.int32), city.to(torch.int16), status.to(torch.uint8), fare]
        for ci, v in enumerate(cols):
            ok = torch.rand(n, generator=g, device=device) >= null_rate
            v = torch.where(ok, v, torch.zeros_like(v))
            w = np_bytes[ci]
            bufs[ci][values_off + c0 * w: values_off + (c0 + n) * w] = v.contiguous().view(torch.uint8)
            pad = (-n) % 8
            okp = torch.cat([ok, torch.zeros(pad, dtype=torch.bool, device=device)]) if pad else ok
            packed = (okp.view(-1, 8).to(torch.uint8) * weights).sum(dim=1, dtype=torch.int32).to(torch.uint8)
            bufs[ci][c0 // 8: c0 // 8 + packed.numel()] = packed  # chunk is a multiple of 8
    return bufs, values_off
