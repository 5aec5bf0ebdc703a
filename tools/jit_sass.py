#!/usr/bin/env python
"""CPU-side look at the NVRTC-specialised kernel of a bench workload: generates + compiles it (AresJitDryRun, no GPU),
disassembles the cubin and prints the instruction mix of the hot loop (the code between the `full` mbarrier wait of a
tile and the `empty` arrive).  Usage: python tools/jit_sass.py [cfg3|cfg3_count|cfg2|cfg4|cfg4_hll] [--no-zone-maps] [--dump DIR]"""
import collections
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import bench
    import test_jit_codegen as TJ
    from aresdb_b200 import cabi as A, synth
    name = next((a for a in sys.argv[1:] if not a.startswith("-")), "cfg3")
    zm = "--no-zone-maps" not in sys.argv
    dump = tempfile.mkdtemp(prefix="aresjit_")
    if "--dump" in sys.argv:
        dump = sys.argv[sys.argv.index("--dump") + 1]
        os.makedirs(dump, exist_ok=True)
    os.environ["ARESDB_B200_JIT_DUMP_DIR"] = dump
    wl = bench.WORKLOADS[name]
    q = wl["query"]()
    rows = wl["rows"] // wl["batches"]
    size, src = TJ._dry_run(A.load_engine(), q, rows=rows, expected_groups=wl["expected_groups"],
                            ranges=synth.zone_map_of_day(0) if zm else None)
    cubin = sorted(Path(dump).glob("*.cubin"))[-1]
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(cubin)], capture_output=True, text=True)
    print((res.stdout + res.stderr).strip().splitlines()[-1])
    lines = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = []
    for l in lines:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops.append((int(m.group(1), 16), m.group(2), l.strip()))
    # hot loop: from the first SYNCS...TRYWAIT after the producer branch to the matching ARRIVE; print the whole mix and
    # the mix between the consumer's try_wait and arrive
    tw = [i for i, o in enumerate(ops) if o[1].startswith("SYNCS.PHASECHK") or "TRYWAIT" in o[2]]
    ar = [i for i, o in enumerate(ops) if o[1].startswith("SYNCS.ARRIVE") and "TRANS" not in o[2]]
    print(f"{len(ops)} SASS instructions; try_wait at {tw[:6]}, arrive at {ar[:6]}")
    def mix(a, b, title):
        c = collections.Counter(o[1].split(".")[0] for o in ops[a:b])
        print(f"--- {title}: {b - a} instructions")
        print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
    if tw and ar:
        # consumer loop body = last try_wait before the last plain arrive ... that arrive
        end = ar[-1] if len(ar) else len(ops)
        cands = [t for t in tw if t < end]
        start = cands[-1] if cands else 0
        mix(start, end + 1, "consumer tile body (one quad of 4 rows per thread)")
    mix(0, len(ops), "whole kernel")
    print("cubin:", cubin, " source:", str(cubin).replace(".cubin", ".cu"))


if __name__ == "__main__":
    main()
