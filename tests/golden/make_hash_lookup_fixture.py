"""Extracts the cuckoo-index golden vectors of the reference's HashLookupTest.CheckLookup / CheckUUID
(query/algorithm_unittest.cu:731-880: indexes created by the reference's golang memstore) into
tests/golden/hash_lookup.npz.  Run in the container that has /root/reference."""
import re
from pathlib import Path

import numpy as np

src = Path("/root/reference/query/algorithm_unittest.cu").read_text()


def grab(test):
    a = src.index(f"TEST(HashLookupTest, {test})")
    m = re.search(r"uint8_t bucketsH\[(\d+)\] = \{([^}]*)\}", src[a:])
    raw = np.array([int(x) for x in re.findall(r"\d+", m.group(2))], np.uint8)
    assert raw.size == int(m.group(1))
    s = re.search(r"\{\(uint32_t\) (\d+), \(uint32_t\) (\d+), \(uint32_t\) (\d+),\s*\(uint32_t\) (\d+)\},\s*(\d+),\s*(\d+),\s*(\d+)\}", src[a:])
    seeds = np.array([int(s.group(i)) for i in range(1, 5)], np.uint32)
    return raw, seeds, np.array([int(s.group(5)), int(s.group(6)), int(s.group(7))], np.int32)


l, ls, lp = grab("CheckLookup")
u, us, up = grab("CheckUUID")
np.savez(Path(__file__).parent / "hash_lookup.npz", lookup_buckets=l, lookup_seeds=ls, lookup_params=lp,
         uuid_buckets=u, uuid_seeds=us, uuid_params=up)
print(l.size, ls, lp, u.size, us, up)
