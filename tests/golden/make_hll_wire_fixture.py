#!/usr/bin/env python
"""Copies the reference's own golden buffers of the application/hll wire format (testing/data/query/hll,
hll_query_results, hll_empty_results — the inputs of query/common/hll_test.go) into tests/golden/hll_wire_format.npz.
Run in the build container (needs /root/reference); the fixture travels, the reference tree does not."""
from pathlib import Path

import numpy as np

REF = Path("/root/reference/testing/data/query")
OUT = Path(__file__).resolve().parent / "hll_wire_format.npz"
np.savez_compressed(OUT, **{name: np.frombuffer((REF / name).read_bytes(), np.uint8) for name in ("hll", "hll_query_results", "hll_empty_results")})
print(OUT, OUT.stat().st_size)
