"""The drop-in boundary itself, without a GPU: both libraries load, export every function the headers under
include/aresdb_b200/ declare (= the reference's cgo surface, query/time_series_aggregate.h:431-621 and
cgoutils/memory.h:51-99, plus the additive plan API), and report failures the way the Go side expects
(cgoutils/utils.go:25-33: a malloc'd C string in CGoCallResHandle.pStrErr that the caller frees; nothing thrown).
No compute call is made here; on a box without a CUDA device every device entry point must fail LOUDLY — the engine
has no CPU fallback."""
import ctypes as C
import re
from pathlib import Path

import pytest

from aresdb_b200 import cabi as A

ROOT = Path(__file__).resolve().parent.parent
HEADERS = sorted((ROOT / "include" / "aresdb_b200").glob("*.h"))
# a declaration: return type, name, '(' at the start of a line (the headers are plain C)
DECL = re.compile(r"^(?:CGoCallResHandle|DeviceMemoryFlags|unsigned long long|uint32_t|int|void)\s+\*?([A-Za-z_]\w*)\s*\(", re.M)


def declared_functions():
    names = {}
    for h in HEADERS:
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        for m in DECL.finditer(text):
            names[m.group(1)] = h.name
    return names


def test_headers_declare_the_reference_surface():
    names = declared_functions()
    for sym in A.ALGORITHM_SYMBOLS + A.PLAN_SYMBOLS + A.MEMORY_SYMBOLS:
        assert sym in names, f"{sym} is bound by cabi.py but not declared in include/aresdb_b200/*.h"
    assert len(A.ALGORITHM_SYMBOLS) == 14       # query/time_series_aggregate.h:431-621


def test_every_declared_function_is_exported():
    lib = A.load_engine()
    missing = []
    for name, header in declared_functions().items():
        if not any(hasattr(dll, name) for dll in (lib.alg, lib.mem)):
            missing.append(f"{name} ({header})")
    assert not missing, "declared but not exported: " + ", ".join(missing)
    # same split as the reference: memory / stream symbols live in libmem, libalgorithm links it
    assert all(hasattr(lib.mem, s) for s in A.MEMORY_SYMBOLS)
    assert all(hasattr(lib.alg, s) for s in A.ALGORITHM_SYMBOLS + A.PLAN_SYMBOLS)


def test_flags_announce_a_device_build_with_hash_reduction():
    lib = A.load_engine()
    # DEVICE_MEMORY_IMPLEMENTATION_FLAG (1) | HASH_REDUCTION_SUPPORT (4); POOLED_MEMORY_FLAG (2) is deliberately not set
    assert lib.get_flags() == 0x1 | 0x4


def _raw(lib, name):
    fn = getattr(lib.alg, name)
    fn.restype = A.CGoCallResHandle
    return fn


def _take_error(h) -> str:
    assert h.pStrErr, "expected an error string"
    msg = C.string_at(h.pStrErr).decode(errors="replace")
    C.CDLL(None).free(C.c_void_p(h.pStrErr))    # the caller owns it (strdup / malloc)
    return msg


def test_out_of_scope_entry_points_answer_with_an_error_string():
    """Joins, geo and non-aggregate expansion are outside this engine: the symbols exist (the Go side links them) and
    return an error string instead of computing anything."""
    lib = A.load_engine()
    h = _raw(lib, "Expand")(A.DimensionVector(), A.DimensionVector(), None, None, 0, 0, None, 0)
    assert _take_error(h)
    h = _raw(lib, "WriteGeoShapeDim")(0, A.DimensionOutputVector(), 0, None, None, 0)
    assert _take_error(h)


def test_invalid_plans_are_rejected_with_a_message_not_a_crash():
    lib = A.load_engine()
    fn = lib.alg.AresJitDryRun
    fn.argtypes = [A.AggSpec, C.POINTER(A.BatchPlan), C.POINTER(C.c_char_p)]
    fn.restype = A.CGoCallResHandle
    spec = A.AggSpec()
    spec.NumDimsPerDimWidth[2] = 1
    spec.AggFunc, spec.MeasureDataType, spec.ReduceMode = A.AGGR_SUM_UNSIGNED, A.Uint32, A.ARES_REDUCE_SORT
    plan = A.BatchPlan()
    plan.NumInsts = 0                                   # no instructions at all
    assert "instruction count" in _take_error(fn(spec, C.byref(plan), None))
    plan.NumInsts, plan.NumColumns = 1, 40              # more columns than the fused path stages
    assert "columns" in _take_error(fn(spec, C.byref(plan), None))
    spec.ReduceMode = 7
    plan.NumColumns = 0
    assert "ReduceMode" in _take_error(fn(spec, C.byref(plan), None))


def test_no_cpu_fallback_without_a_device():
    """On a box without a CUDA device a device entry point fails with the CUDA error text; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = A.load_engine()
    with pytest.raises(A.AresError):
        lib.AggStateCreate(A.AggSpec(), None, 0)
    with pytest.raises(A.AresError):
        lib.InitIndexVector(None, 0, 16, None, 0)
