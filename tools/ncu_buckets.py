#!/usr/bin/env python
"""Per-instruction view of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv):
prints instructions executed per warp pass, average active threads and stall samples in buckets
of N SASS instructions, or every instruction of a range.

  python tools/ncu_buckets.py src.csv <warp_passes> [bucket] [lo hi]
"""
import csv
import sys


def main():
    path, passes = sys.argv[1], float(sys.argv[2])
    bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rng = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else None
    rows = list(csv.reader(open(path)))
    h = rows[1]
    body = rows[2:]
    isrc, iex, ith, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    tot = sum(float(x[iex]) for x in body)
    smp = sum(float(x[ismp]) for x in body)
    print(f"{len(body)} SASS instructions, {tot:.4g} executed = {tot / passes:.1f} per warp pass, {smp:.0f} samples")
    if rng:
        for i in range(rng[0], min(rng[1], len(body))):
            x = body[i]
            e = float(x[iex])
            print(f"{i:5d} {e / passes:6.2f} thr {float(x[ith]) / max(e, 1):5.1f} smp {float(x[ismp]) / smp * 100:5.2f}%  {x[isrc][:90]}")
        return
    for s in range(0, len(body), bucket):
        blk = body[s:s + bucket]
        e = sum(float(x[iex]) for x in blk)
        t = sum(float(x[ith]) for x in blk)
        m = sum(float(x[ismp]) for x in blk)
        print(f"{s:5d} {e / passes:6.1f}/pass thr {t / max(e, 1):5.1f} smp {m / smp * 100:5.1f}%  {blk[0][isrc][:70]}")


if __name__ == "__main__":
    main()
