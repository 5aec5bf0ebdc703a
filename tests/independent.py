"""An INDEPENDENT restatement of the synthetic benchmark workloads in plain torch ops (boolean masks,
integer division, float64 / int64 index_add, scatter-max) over the same column buffers the engine
reads.  TEST INFRASTRUCTURE: used by tests/ (pinned against the oracle at small sizes on the CPU, then
trusted at BASELINE sizes on the GPU) and by bench.py's one-off self-verification outside the timed
region.  Nothing under aresdb_b200/ imports it; it shares no code with the engine or the oracle.

Column buffers are the [null bitmap][values] byte tensors of synth.generate_batch_cuda (mode 2).
Group space: time index (bucket ordinal from BASE_TS, one extra index for NULL) x city index (0..127,
one extra for NULL); a NULL dimension carries value 0 with validity 0 (synth stores 0 under a NULL).
"""
from __future__ import annotations

import numpy as np

CITY_SPACE = 129          # city values 0..127 + NULL
HLL_REGS = 1 << 14


def decode_columns(bufs, values_off: int, rows: int):
    """(ts i64, city i64, status i64, fare f32), (valid bool x4) views / copies of one batch."""
    import torch
    ts = bufs[0][values_off:values_off + 4 * rows].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    city = bufs[1][values_off:values_off + 2 * rows].view(torch.int16).to(torch.int64) & 0xFFFF
    status = bufs[2][values_off:values_off + rows].to(torch.int64)
    fare = bufs[3][values_off:values_off + 4 * rows].view(torch.float32)
    valid = []
    shifts = torch.arange(8, device=bufs[0].device, dtype=torch.uint8)
    for b in bufs:
        bits = b[: (rows + 7) // 8]
        valid.append(((bits.unsqueeze(1) >> shifts) & 1).reshape(-1)[:rows].bool())
    return (ts, city, status, fare), valid


# name -> (filters, time bucket seconds or None, city is a dimension, measure kind)
WORKLOADS = {
    "cfg3": (("status", "fare", "city", "time"), 3600, True, "sum"),
    "cfg3_count": (("status", "fare", "city"), 3600, True, "count"),
    "cfg2": (("status",), None, True, "sum"),
    "cfg4": ((), 60, True, "sum"),
    "cfg4_hll": (("status",), 86400, True, "hll"),
}


def _mm3_fmix(k):
    def lsr(x, r):
        return (x >> r) & ((1 << (64 - r)) - 1)
    k = k ^ lsr(k, 33)
    k = k * -49064778989728563            # 0xff51afd7ed558ccd as int64
    k = k ^ lsr(k, 33)
    k = k * -4265267296055464877          # 0xc4ceb9fe1a85ec53
    return k ^ lsr(k, 33)


def murmur3_x64_128_lo_u32(v):
    """Low 64 bits of MurmurHash3_x64_128(seed 0) of the 4 little-endian bytes of each value (int64 tensor
    holding uint32 values); int64 arithmetic wraps like uint64."""
    c1, c2 = -8663945395140668459, 5545529020109919103   # 0x87c37b91114253d5, 0x4cf5ad432745937f
    k1 = v * c1
    k1 = (k1 << 31) | ((k1 >> 33) & ((1 << 31) - 1))
    k1 = k1 * c2
    h1 = k1 ^ 4
    h2 = h1 * 0 + 4
    h1 = h1 + h2
    h2 = h2 + h1
    h1 = _mm3_fmix(h1)
    h2 = _mm3_fmix(h2)
    return h1 + h2


def hll_value(ts):
    """(register, rho + 1) of GetHLLValue(uint32) with the CUDA semantics of the int shift (bits 14..31
    decide; none set: rho = 50) — reference query/functor.hpp:431-466."""
    import torch
    h = murmur3_x64_128_lo_u32(ts)
    reg = h & (HLL_REGS - 1)
    x = (h >> 14) & 0x3FFFF
    lsb = x & (-x)
    # count of trailing zeros = exponent of the (exactly representable) lowest set bit; no log2: its CUDA form is not
    # exact on powers of two
    ctz = (lsb.clamp(min=1).to(torch.float32).view(torch.int32) >> 23).to(torch.int64) - 127
    rho = torch.where(x == 0, torch.full_like(x, 50), ctz)
    return reg, rho + 1


class Expected:
    """Accumulates batches of one workload; `present` / `vals` (and `regs` for hll) over the group space."""

    def __init__(self, name: str, num_days: int, device, base_ts: int, time_lo: int | None = None, time_hi: int | None = None):
        import torch
        self.torch, self.name, self.base_ts = torch, name, base_ts
        self.filters, self.step, self.city_dim, self.kind = WORKLOADS[name]
        self.time_lo, self.time_hi = time_lo, time_hi
        self.tn = (num_days * 86400 // self.step + 1) if self.step else 1
        n = self.tn * CITY_SPACE
        self.present = torch.zeros(n, dtype=torch.bool, device=device)
        self.vals = torch.zeros(n, dtype=torch.float64 if self.kind == "sum" else torch.int64, device=device)
        self.regs = {}   # hll: group index -> uint8[16384], filled by finish()
        self._hll = torch.zeros(0, dtype=torch.int64, device=device) if self.kind == "hll" else None
        self.rows_kept = 0

    def add_batch(self, bufs, values_off: int, rows: int, chunk: int = 1 << 25):
        torch = self.torch
        for c0 in range(0, rows, chunk):   # bounded temporaries
            n = min(chunk, rows - c0)
            assert c0 % 8 == 0
            (ts, city, status, fare), (vts, vcity, vstatus, vfare) = _decode_range(bufs, values_off, c0, n)
            keep = torch.ones(n, dtype=torch.bool, device=ts.device)
            if "status" in self.filters:
                keep &= vstatus & (status == 1)
            if "fare" in self.filters:
                keep &= vfare & (fare > 5.0)
            if "city" in self.filters:
                keep &= vcity & (city != 0)
            if "time" in self.filters:
                keep &= vts & (ts >= self.time_lo) & (ts < self.time_hi)
            if self.step:
                tidx = torch.where(vts, (ts - ts % self.step - self.base_ts) // self.step, torch.full_like(ts, self.tn - 1))
                assert int(tidx.min()) >= 0 and int(tidx[vts].max() if vts.any() else 0) < self.tn - 1
            else:
                tidx = torch.zeros_like(ts)
            cidx = torch.where(vcity, city, torch.full_like(city, CITY_SPACE - 1))
            assert int(city.max()) < CITY_SPACE - 1
            g = (tidx * CITY_SPACE + cidx)[keep]
            self.rows_kept += int(g.numel())
            self.present[g] = True
            if self.kind == "sum":     # NULL fare -> the identity 0.0
                contrib = torch.where(vfare, fare.double(), torch.zeros((), dtype=torch.float64, device=ts.device))[keep]
                self.vals.index_add_(0, g, contrib)
            elif self.kind == "count":
                self.vals.index_add_(0, g, torch.ones_like(g))
            else:                       # hll of request_at; a NULL measure is the identity value 0 (register 0, rho 0)
                reg, rho1 = hll_value(ts)
                reg = torch.where(vts, reg, torch.zeros_like(reg))[keep]
                rho1 = torch.where(vts, rho1, torch.ones_like(rho1))[keep]
                key = (g * HLL_REGS + reg) * 64 + rho1   # max rho per (group, register) = max key per (group, register)
                self._hll = torch.unique(torch.cat([self._hll, torch.unique(key)]))

    def hll_registers(self) -> dict:
        """group index -> uint8[16384] of rho + 1."""
        key = self._hll.cpu().numpy()
        gr, rho1 = key // 64, (key % 64).astype(np.uint8)
        out = {}
        order = np.argsort(gr, kind="stable")            # ascending key: the last entry of a (group, register) run is its max
        gr, rho1 = gr[order], rho1[order]
        last = np.r_[gr[1:] != gr[:-1], True]
        for k, r in zip(gr[last], rho1[last]):
            g, reg = divmod(int(k), HLL_REGS)
            out.setdefault(g, np.zeros(HLL_REGS, np.uint8))[reg] = r
        return out

    # ---- comparison with an engine result --------------------------------------------------------
    def group_index(self, res):
        """Group-space index of every row of a QueryResult (dimension values AND validity bytes are checked
        for consistency: a NULL dimension must carry value 0)."""
        q = res.query
        n = res.groups
        tidx = np.zeros(n, np.int64)
        cidx = np.zeros(n, np.int64)
        for qi, dt in enumerate(q.dim_types):
            vals = res.dim_values[qi].reshape(n, -1)
            valid = res.dim_valid[qi].astype(bool)
            if vals.shape[1] == 4:     # the time bucket
                v = vals.copy().view(np.uint32).reshape(-1).astype(np.int64)
                assert (v[~valid] == 0).all(), "NULL time dimension with a non-zero value"
                assert ((v[valid] - self.base_ts) % self.step == 0).all(), "time dimension is not a bucket start"
                tidx = np.where(valid, (v - self.base_ts) // self.step, self.tn - 1)
            else:
                v = vals.copy().view(np.uint16).reshape(-1).astype(np.int64)
                assert (v[~valid] == 0).all(), "NULL city dimension with a non-zero value"
                cidx = np.where(valid, v, CITY_SPACE - 1)
        assert (tidx >= 0).all() and (tidx < self.tn).all() and (cidx < CITY_SPACE).all(), "dimension value outside the table"
        return tidx * CITY_SPACE + cidx

    def check(self, res) -> dict:
        """Raises AssertionError on any difference; returns a small summary."""
        idx = self.group_index(res)
        present = self.present.cpu().numpy()
        assert len(np.unique(idx)) == res.groups, "duplicate groups in the result"
        assert res.groups == int(present.sum()), f"{res.groups} groups, expected {int(present.sum())}"
        assert present[idx].all(), "a group of the result has no surviving row"
        exp = self.vals.cpu().numpy()[idx]
        got = res.measures
        if self.kind == "sum":
            assert got.dtype == np.float64
            bad = np.nonzero(got.view(np.uint64) != exp.view(np.uint64))[0]
            assert bad.size == 0, (f"{bad.size} double sums differ (they are exact on quantised fares); first: group {idx[bad[0]]} "
                                   f"got {got[bad[0]]!r} expected {exp[bad[0]]!r}")
        else:
            assert (got.astype(np.int64) == exp).all(), "counts differ"
            assert int(got.astype(np.int64).sum()) == self.rows_kept
        return {"groups": int(res.groups), "rows_kept": self.rows_kept}

    def check_hll(self, hres) -> dict:
        idx = self.group_index(hres.dims)
        present = self.present.cpu().numpy()
        assert hres.groups == int(present.sum()) and present[idx].all() and len(np.unique(idx)) == hres.groups
        exp = self.hll_registers()
        got = hres.dense_registers()
        for g, row in zip(idx.tolist(), hres.dims.rows):
            assert (got[row] == exp[g]).all(), f"registers of group {g} differ"
        return {"groups": int(hres.groups), "rows_kept": self.rows_kept}


def _decode_range(bufs, values_off: int, c0: int, n: int):
    """decode_columns of rows [c0, c0 + n) (c0 a multiple of 8)."""
    import torch
    ts = bufs[0][values_off + 4 * c0:values_off + 4 * (c0 + n)].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    city = bufs[1][values_off + 2 * c0:values_off + 2 * (c0 + n)].view(torch.int16).to(torch.int64) & 0xFFFF
    status = bufs[2][values_off + c0:values_off + c0 + n].to(torch.int64)
    fare = bufs[3][values_off + 4 * c0:values_off + 4 * (c0 + n)].view(torch.float32)
    shifts = torch.arange(8, device=bufs[0].device, dtype=torch.uint8)
    valid = []
    for b in bufs:
        bits = b[c0 // 8: (c0 + n + 7) // 8]
        valid.append(((bits.unsqueeze(1) >> shifts) & 1).reshape(-1)[:n].bool())
    return (ts, city, status, fare), valid


def host_batch_buffers(hb, device="cpu"):
    """The [null bitmap][values] byte tensors of a synth.HostBatch (what generate_batch_cuda lays out)."""
    import torch
    rows = hb.num_rows
    values_off = ((rows + 7) // 8 + 1 + 63) // 64 * 64
    bufs = []
    for v, ok in zip(hb.values, hb.valid):
        raw = np.zeros(values_off + (v.nbytes + 64) // 64 * 64, np.uint8)
        bits = np.packbits(ok != 0, bitorder="little")
        raw[:bits.size] = bits
        raw[values_off:values_off + v.nbytes] = np.ascontiguousarray(v).view(np.uint8)
        bufs.append(torch.from_numpy(raw).to(device))
    return bufs, values_off
