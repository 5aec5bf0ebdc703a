#!/usr/bin/env python
"""Evidence for the drop-in path: times the per-node entry points the UNCHANGED Go driver would call — Sort, Reduce,
HashReduce, HyperLogLog (reference query/sort_reduce.cu, hash_reduction.cu, hll.cu) — on `--rows` rows (default 1e8) of a
cfg3-shaped dimension block (u32 hour bucket x u16 city + validity bytes, f64 measures) with CUDA events, and prints one
JSON line with ms and the effective HBM bandwidth against the bytes each call must at least move.
    python tools/legacy_bench.py [--rows N] [--groups G] [--reps R]"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--groups", type=int, default=19200)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from aresdb_b200 import cabi as A
    from aresdb_b200.executor import dim_offsets
    lib = A.load_engine()
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    n, nd = args.rows, (0, 0, 1, 1, 0)
    offs, nulls, widths, total = dim_offsets(nd, n)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    hours = max(args.groups // 100, 1)
    block = torch.zeros(total, dtype=torch.uint8, device=dev)
    block[offs[0]:offs[0] + 4 * n] = (torch.randint(0, hours, (n,), generator=g, device=dev, dtype=torch.int32) * 3600 + 1_726_963_200).view(torch.uint8)
    block[offs[1]:offs[1] + 2 * n] = torch.randint(1, 101, (n,), generator=g, device=dev, dtype=torch.int16).view(torch.uint8)
    block[nulls[0]:nulls[0] + n] = 1
    block[nulls[1]:nulls[1] + n] = 1
    meas = (torch.randint(0, 6400, (n,), generator=g, device=dev, dtype=torch.int32).double() / 64.0)
    hashv = torch.zeros(n, dtype=torch.int64, device=dev)
    index = torch.arange(n, dtype=torch.int32, device=dev)
    oblock, ohash, oindex = torch.zeros_like(block), torch.zeros_like(hashv), torch.zeros_like(index)
    omeas = torch.zeros_like(meas)
    kin = A.make_dimension_vector(block.data_ptr(), hashv.data_ptr(), index.data_ptr(), nd, n)
    kout = A.make_dimension_vector(oblock.data_ptr(), ohash.data_ptr(), oindex.data_ptr(), nd, n)
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=args.reps, setup=None):
        best = []
        for _ in range(reps + 1):
            if setup:
                setup()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn()
            e.record()
            torch.cuda.synchronize()
            best.append(s.elapsed_time(e))
        return float(np.median(best[1:])), r

    out = {"rows": n, "groups": args.groups}
    l0 = lib.kernel_launch_count()
    ms, _ = timed(lambda: lib.Sort(kin, n, stream, 0), setup=lambda: index.copy_(torch.arange(n, dtype=torch.int32, device=dev)))
    out["Sort"] = {"ms": ms, "launches": (lib.kernel_launch_count() - l0) // (args.reps + 1),
                   "min_bytes": n * (8 + 12 + 12), "GBps_vs_min": n * 32 / ms / 1e6,
                   "note": "hash 8 B/row read + one read and one write of (u64 hash, u32 index); the LSD sort makes 8 passes of that"}
    ms, gr = timed(lambda: lib.Reduce(kin, meas.data_ptr(), kout, omeas.data_ptr(), 8, n, A.AGGR_SUM_FLOAT, stream, 0))
    out["Reduce"] = {"ms": ms, "groups": gr, "min_bytes": n * (8 + 4 + 8), "GBps_vs_min": n * 20 / ms / 1e6}
    ms, gh = timed(lambda: lib.HashReduce(A.make_dimension_vector(block.data_ptr(), None, None, nd, n), meas.data_ptr(),
                                          A.make_dimension_vector(oblock.data_ptr(), None, None, nd, n), omeas.data_ptr(), 8, n,
                                          A.AGGR_SUM_FLOAT, stream, 0))
    out["HashReduce"] = {"ms": ms, "groups": gh, "min_bytes": n * (8 + 8), "GBps_vs_min": n * 16 / ms / 1e6}
    # HyperLogLog: one batch, last = true; values = rho << 16 | reg of a random stream
    vals = torch.randint(0, 1 << 14, (n,), generator=g, device=dev, dtype=torch.int32) | (torch.randint(0, 20, (n,), generator=g, device=dev, dtype=torch.int32) << 16)
    prev_vals = torch.zeros(1, dtype=torch.int32, device=dev)
    vec, size, counts = C.c_void_p(), C.c_size_t(), C.c_void_p()

    def hll():
        index.copy_(torch.arange(n, dtype=torch.int32, device=dev))
        r = lib.HyperLogLog(kin, kout, prev_vals.data_ptr(), vals.data_ptr(), 0, n, True, C.byref(vec), C.byref(size), C.byref(counts), stream, 0)
        for p in (vec, counts):
            if p.value:
                lib.DeviceFree(p, 0)
        return r

    ms, gd = timed(hll, reps=max(1, args.reps - 1))
    out["HyperLogLog"] = {"ms": ms, "dims": gd, "min_bytes": n * (8 + 4 + 4), "GBps_vs_min": n * 16 / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
