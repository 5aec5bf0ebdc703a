"""The boundary, checked against the REFERENCE's own headers by the C compiler (needs /root/reference; skipped on the GPU box):
tests/abi/abi_probe.c is compiled once against query/time_series_aggregate.h + cgoutils/memory.h and once against
include/aresdb_b200/*.h — every struct size, field offset and enum value must agree — and the reference-header build is
linked against aresdb_b200/lib and run, so a translation unit that only ever saw the reference's prototypes binds to the
B200 libraries unchanged."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def _build(tmp, name, flags, link):
    exe = tmp / name
    cmd = ["gcc", "-std=gnu11", "-O0", "-w", str(ROOT / "tests" / "abi" / "abi_probe.c"), "-o", str(exe)] + flags
    if link:
        lib = ROOT / "aresdb_b200" / "lib"
        cmd += [f"-L{lib}", "-lalgorithm", "-lmem", f"-Wl,-rpath,{lib}"]
    else:
        cmd += ["-Wl,--unresolved-symbols=ignore-all"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.skipif(not (REF / "query" / "time_series_aggregate.h").exists(), reason="reference tree not present")
def test_struct_layouts_and_enums_equal_the_reference_headers(tmp_path):
    ours = _build(tmp_path, "probe_ours", [f"-I{ROOT / 'include'}"], link=False)
    ref = _build(tmp_path, "probe_ref", ["-DUSE_REFERENCE", f"-I{REF}"], link=False)
    a = subprocess.run([str(ours)], capture_output=True, text=True).stdout
    b = subprocess.run([str(ref)], capture_output=True, text=True).stdout
    assert a.count("\n") > 130
    assert a == b


@pytest.mark.skipif(not (REF / "query" / "time_series_aggregate.h").exists(), reason="reference tree not present")
def test_reference_header_build_links_against_the_b200_libraries(tmp_path):
    lib = ROOT / "aresdb_b200" / "lib" / "libalgorithm.so"
    if not lib.exists():
        pytest.skip("engine not built")
    exe = _build(tmp_path, "probe_link", ["-DUSE_REFERENCE", f"-I{REF}"], link=True)
    r = subprocess.run([str(exe), "link"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "link symbols 10" in r.stdout and "link GetFlags" in r.stdout
