"""Multi-GPU execution of one aggregate query: table shards / archive batches are independent units
(the reference processes them one after the other and only couples them through the carried result
vectors), so batch i goes to rank i mod N, every rank aggregates its batches into a
private group table, and ONE exchange step merges the per-rank results: an all-gather of the compact
(dimension block, measure vector) pairs over NCCL followed by a local re-aggregation on every rank
with the aggregate's combine rule — what the reference's broker does with JSON results
(broker/result_merge.go:80-105: sum/count add, min/max).  Results are identical on every rank.
"""
from __future__ import annotations

import numpy as np

from . import cabi as A
from .executor import _ResultBuffers, dim_offsets
from .query import AggQuery, QueryResult


def assign_batches(num_batches: int, world: int, rank: int) -> list[int]:
    """Round-robin placement of batch ids."""
    return [b for b in range(num_batches) if b % world == rank]


def pad_result(space, q: AggQuery, src: _ResultBuffers, groups: int, capacity: int) -> _ResultBuffers:
    """Re-lays a result block out for `capacity` rows (all ranks must gather equal-sized tensors)."""
    pad = _ResultBuffers(space, q, max(capacity, 1))
    so, sn, widths, _ = dim_offsets(q.num_dims_per_width, src.capacity)
    do, dn, _, _ = dim_offsets(q.num_dims_per_width, pad.capacity)
    for p, w in enumerate(widths):
        space.copy(pad.dims, do[p], src.dims, so[p], w * groups)
        space.copy(pad.dims, dn[p], src.dims, sn[p], groups)
    space.copy(pad.measures, 0, src.measures, 0, q.measure_bytes * groups)
    return pad


class ShardedFusedQuery:
    """One rank's half of a sharded query on the B200 engine (torch.distributed / NCCL plumbing)."""

    def __init__(self, lib, space, q: AggQuery, expected_groups: int = 0, merged_groups: int | None = None):
        from .executor import FusedBatchExecutor
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.lib, self.space, self.q = lib, space, q
        self.local = FusedBatchExecutor(lib, space, q, expected_groups)
        # the merged state sees every rank's groups: same table hint (and, for hll, the same table mode)
        merged_groups = expected_groups if merged_groups is None else merged_groups
        self.merged = FusedBatchExecutor(lib, space, q, merged_groups) if self.world > 1 else None
        # Fixed-capacity exchange (queries that do not announce more groups than this): every rank sends
        # [row count | dimension block | measures] for EXCHANGE_ROWS rows in ONE all-gather, so no count has to be agreed on
        # first (one collective and one host sync less per query); a rank with more rows flags it in its header and
        # every rank repeats the step with the exact-size protocol.
        import os
        self._fixed_cap = self.EXCHANGE_ROWS if (self.world > 1 and expected_groups <= self.EXCHANGE_ROWS and not q.is_hll
                                                 and os.environ.get("ARESDB_B200_EXCHANGE", "fixed") != "exact") else 0
        self._send = self._recv = None
        self._peer, self._peer_ok = None, None   # exchange over peer memory: set up at the first exchange

    def reset(self):
        self.local.reset()

    def process_batch(self, batch, stream=None):
        self.local.process_batch(batch, stream)

    EXCHANGE_ROWS = 32768   # rows of a fixed part (= what the engine's single-launch export handles)
    _HDR = 64

    def _exchange_fixed(self):
        """Device-only exchange: AggStateExportPart (one launch, the row count stays on the device) -> ONE all-gather of
        the fixed parts -> AggStateMergeParts (one launch over all parts).  The host waits for nothing here; a rank with
        more rows than a part holds marks its header and the merged state's finalize reports it (-> exact protocol)."""
        import torch
        q, sp, dist, lib = self.q, self.space, self.dist, self.lib
        cap = self._fixed_cap
        _, _, _, dim_bytes = dim_offsets(q.num_dims_per_width, cap)
        dim_bytes = (dim_bytes + 15) // 16 * 16
        part = (self._HDR + dim_bytes + q.measure_bytes * cap + 63) // 64 * 64
        if self._send is None:
            self._send = torch.zeros(part, dtype=torch.uint8, device=sp.dev)
            self._recv = torch.empty(self.world * part, dtype=torch.uint8, device=sp.dev)
        if self._peer is None and self._peer_ok is None:
            self._setup_peers(part)
        if self._peer is not None:
            return self._exchange_peers(cap, dim_bytes)
        lib.AggStateExportPart(self.local.state, self._send.data_ptr(), cap, self._HDR, self._HDR + dim_bytes, sp.stream, sp.device)
        dist.all_gather_into_tensor(self._recv, self._send)
        self.merged.reset()
        lib.AggStateMergeParts(self.merged.state, self._recv.data_ptr(), self.world, part, cap, self._HDR, self._HDR + dim_bytes,
                               sp.stream, sp.device)
        return self.world * cap

    _FLAGS = 256   # two parities x 16 ranks x uint32, in front of the two receive buffers

    def _setup_peers(self, part: int):
        """Maps every rank's receive buffer into every process (torch symmetric memory: CUDA IPC / fabric handles over
        NVLink) for the exchange over peer memory.  All ranks agree on the outcome; on failure the NCCL all-gather stays."""
        import os
        import sys
        import torch
        ok = 0
        if os.environ.get("ARESDB_B200_EXCHANGE", "peer") == "peer" and self.world <= 16:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(self._FLAGS + 2 * self.world * part, dtype=torch.uint8, device=self.space.dev)
                buf.zero_()
                hdl = symm.rendezvous(buf, self.dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                torch.cuda.synchronize()
                self._peer = dict(buf=buf, hdl=hdl, ptrs=ptrs, epoch=0, part=part)
                ok = 1
            except Exception as e:   # no peer access between these GPUs, or a torch without symmetric memory
                print(f"[aresdb_b200] exchange over peer memory unavailable ({type(e).__name__}: {e}); using the NCCL all-gather",
                      file=sys.stderr)
        agree = torch.tensor([ok], device=self.space.dev, dtype=torch.int32)
        self.dist.all_reduce(agree, op=self.dist.ReduceOp.MIN)   # (also: nobody writes before everybody has zeroed its flags)
        self._peer_ok = bool(agree.item())
        if not self._peer_ok:
            self._peer = None

    def _exchange_peers(self, cap: int, dim_bytes: int):
        """AggStateExportPartToPeers (ONE launch: export + copy of the part into every peer's receive buffer over NVLink +
        the arrival flags) -> AggStateMergePartsWhenFlagged (the merge kernel waits for the peers' flags itself).  No
        collective call, no host wait; receive buffers alternate by epoch parity (a rank is at most one exchange ahead)."""
        import ctypes as C
        pe, sp, lib, w, r = self._peer, self.space, self.lib, self.world, self.rank
        pe["epoch"] += 1
        epoch, par, part = pe["epoch"], pe["epoch"] & 1, pe["part"]
        base = self._FLAGS + par * w * part
        slots = (C.c_void_p * w)(*[pe["ptrs"][p] + base + r * part for p in range(w)])
        flags = (C.c_void_p * w)(*[pe["ptrs"][p] + par * 64 + r * 4 for p in range(w)])
        lib.AggStateExportPartToPeers(self.local.state, slots, flags, w, r, part, cap, self._HDR, self._HDR + dim_bytes, epoch,
                                      sp.stream, sp.device)
        self.merged.reset()
        lib.AggStateMergePartsWhenFlagged(self.merged.state, pe["ptrs"][r] + base, w, part, cap, self._HDR, self._HDR + dim_bytes,
                                          pe["ptrs"][r] + par * 64, epoch, sp.stream, sp.device)
        return w * cap

    def _exchange(self):
        if self._fixed_cap:
            return self._exchange_fixed()
        return self._exchange_exact()

    def _exchange_exact(self):
        """The one exchange step: every rank exports its table (AggStateExport: unordered rows, no
        sort), ONE all-gather moves [dim block | measure vector] of every rank, and every rank folds
        all of them into `self.merged`.  Returns the number of rows gathered (an upper bound of the
        merged group count).  For hll queries the rows are the carried (group, register) entries."""
        import torch
        q, sp, dist, lib = self.q, self.space, self.dist, self.lib
        n = self.local.group_count()
        counts = torch.zeros(self.world, dtype=torch.int64, device=sp.dev)
        counts[self.rank] = n
        dist.all_reduce(counts)
        counts = counts.tolist()
        cap = max(max(counts), 1)
        _, _, _, dim_bytes = dim_offsets(q.num_dims_per_width, cap)
        dim_bytes = (dim_bytes + 15) // 16 * 16
        part = dim_bytes + q.measure_bytes * cap
        gathered = torch.empty(self.world * part, dtype=torch.uint8, device=sp.dev)
        mine = gathered[self.rank * part:(self.rank + 1) * part]
        if n:
            dv = A.make_dimension_vector(mine.data_ptr(), None, None, q.num_dims_per_width, cap)
            lib.AggStateExport(self.local.state, dv, mine.data_ptr() + dim_bytes, sp.stream, sp.device)
        dist.all_gather_into_tensor(gathered, mine.clone())
        self.merged.reset()
        base = gathered.data_ptr()
        for r in range(self.world):
            if counts[r]:
                dv = A.make_dimension_vector(base + r * part, None, None, q.num_dims_per_width, cap)
                self.merged.merge(dv, base + r * part + dim_bytes, counts[r])
        self._keep = gathered   # merge is asynchronous: keep the gathered rows alive until finalize
        return int(sum(counts))

    def finalize(self):
        """(groups, result buffers) of the WHOLE query, identical on every rank."""
        if self.world == 1:
            return self.local.finalize_into()
        if self._fixed_cap:
            self._exchange_fixed()
            try:
                return self.merged.finalize_into()
            except A.AresError as e:
                if "exchange part truncated" not in str(e):
                    raise
        return self.merged.finalize_into(self._exchange_exact())

    def finalize_hll(self):
        """hll queries: the HLLResult of the WHOLE query, identical on every rank."""
        if self.world == 1:
            return self.local.hll_result()
        self._exchange()
        return self.merged.hll_result()

    def close(self):
        self.local.close()
        if self.merged:
            self.merged.close()


# ---- host-side mirror (gloo / CPU): same protocol on QueryResults, used by the CPU test-suite ------
_COMBINE = {A.AGGR_SUM_UNSIGNED: np.add, A.AGGR_SUM_SIGNED: np.add, A.AGGR_SUM_FLOAT: np.add,
            A.AGGR_MIN_UNSIGNED: np.minimum, A.AGGR_MIN_SIGNED: np.minimum, A.AGGR_MIN_FLOAT: np.minimum,
            A.AGGR_MAX_UNSIGNED: np.maximum, A.AGGR_MAX_SIGNED: np.maximum, A.AGGR_MAX_FLOAT: np.maximum}


def merge_results_host(q: AggQuery, parts: list[dict]) -> dict:
    """Folds per-rank {packed dim row: measure} maps with the aggregate's combine rule."""
    op = _COMBINE[q.agg_func]
    out: dict = {}
    for part in parts:
        for k, v in part.items():
            out[k] = op(out[k], v) if k in out else v
    return out


def all_gather_results_host(dist, result: QueryResult) -> list[dict]:
    """all_gather_object of the compact result maps (gloo works on CPU tensors / objects)."""
    world = dist.get_world_size()
    gathered = [None] * world
    dist.all_gather_object(gathered, result.as_dict())
    return gathered


def merge_hll_results_host(parts: list[dict]) -> dict:
    """hll queries: per-rank {packed dim row: uint8[16384] registers} maps folded with the per-register
    maximum (broker/result_merge.go: HLL merge = Merge of the register sets)."""
    out: dict = {}
    for part in parts:
        for k, regs in part.items():
            out[k] = np.maximum(out[k], regs) if k in out else regs
    return out


def all_gather_hll_host(dist, result) -> list[dict]:
    """all_gather_object of the dense register maps of an HLLResult."""
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, result.dense_registers())
    return gathered
