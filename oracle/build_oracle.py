"""TEST INFRASTRUCTURE — compiles the plain-C restatement (oracle/aql_oracle.c) into
oracle/build/liboracle.so with gcc.  Loaded only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs."""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent


def build(force: bool = False) -> Path:
    out = HERE / "build" / "liboracle.so"
    src = HERE / "aql_oracle.c"
    deps = [src] + list((ROOT / "include" / "aresdb_b200").glob("*.h"))
    if not force and out.exists() and out.stat().st_mtime > max(p.stat().st_mtime for p in deps):
        return out
    out.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["gcc", "-O3", "-std=gnu11", "-fPIC", "-shared", "-fno-strict-aliasing", "-Wall", "-Wno-unused-function",
           f"-I{ROOT / 'include'}", str(src), "-o", str(out), "-lm"]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
