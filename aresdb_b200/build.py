"""Builds the two C-ABI shared libraries of the B200 engine IN-TREE with nvcc for sm_100a.

    aresdb_b200/lib/libmem.so        memory / stream half of the cgo boundary   (csrc/memory_pool.cu)
    aresdb_b200/lib/libalgorithm.so  query kernels + entry points, links libmem (csrc/*.cu)

Same library names and split as the reference (CMakeLists.txt:144-203) so that the Go side's
`#cgo LDFLAGS: -lalgorithm` / `-lmem` link unchanged.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "aresdb_b200" / "csrc"
LIB = ROOT / "aresdb_b200" / "lib"
OBJ = ROOT / "build" / "obj"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          f"-I{ROOT / 'include'}", f"-I{CSRC}", "--expt-relaxed-constexpr",
          "-Xcudafe", "--diag_suppress=177"]

MEM_SRCS = ["memory_pool.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the B200 engine cannot be built (no CPU fallback exists)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(str(p).encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def _headers():
    return list(CSRC.glob("*.cuh")) + list((ROOT / "include" / "aresdb_b200").glob("*.h"))


def _compile(nvcc: str, src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".sha")
    dig = _digest([src] + _headers())
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [nvcc, *ARCH, *COMMON, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        print(r.stderr)
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, jobs: int | None = None) -> dict:
    nvcc = _nvcc()
    LIB.mkdir(parents=True, exist_ok=True)
    OBJ.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 4)) as ex:
        objs = dict(zip([s.name for s in srcs], ex.map(lambda s: _compile(nvcc, s, verbose), srcs)))
    mem_objs = [str(objs[n]) for n in MEM_SRCS]
    alg_objs = [str(o) for n, o in objs.items() if n not in MEM_SRCS]
    libmem = LIB / "libmem.so"
    libalg = LIB / "libalgorithm.so"
    link = [nvcc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-cudart", "shared"]
    newest = max(Path(o).stat().st_mtime for o in mem_objs + alg_objs)
    if not libmem.exists() or libmem.stat().st_mtime < newest:
        subprocess.run([*link, *mem_objs, "-o", str(libmem), "-Xlinker", "-soname", "-Xlinker", "libmem.so"], check=True)
    if not libalg.exists() or libalg.stat().st_mtime < newest:
        subprocess.run([*link, *alg_objs, "-o", str(libalg), f"-L{LIB}", "-lmem", "-Xlinker", "-soname", "-Xlinker",
                        "libalgorithm.so", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"], check=True)
    return {"libmem": str(libmem), "libalgorithm": str(libalg)}


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv)
    print(out)
