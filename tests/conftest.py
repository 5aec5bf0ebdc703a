"""pytest wiring: the `gpu` marker and the three C-ABI backends every parity test can bind.

  ref     the reference's own QUERY_MODE=HOST build (oracle/_ref, built by oracle/build_ref.sh)
  oracle  the plain-C restatement under oracle/ (built by oracle/build_oracle.py)
  b200    the CUDA engine (aresdb_b200/lib) — needs a GPU, so every use is marked `gpu`
"""
import sys
from pathlib import Path

import os

import pytest

# a specialised kernel that fails to compile must fail the test, not fall back to the interpreter kernel silently
os.environ.setdefault("ARESDB_B200_JIT_STRICT", "1")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _backend_param(name):
    marks = [pytest.mark.gpu] if name == "b200" else []
    return pytest.param(name, marks=marks, id=name)


@pytest.fixture(params=[_backend_param("ref"), _backend_param("oracle"), _backend_param("b200")])
def backend(request):
    """A harness.Backend for each implementation of the C ABI."""
    import harness
    return harness.get_backend(request.param)


@pytest.fixture(params=[_backend_param("oracle"), _backend_param("b200")])
def impl(request):
    """Backends that are compared AGAINST the reference build (never the reference itself)."""
    import harness
    return harness.get_backend(request.param)


@pytest.fixture
def ref():
    import harness
    return harness.get_backend("ref")


@pytest.fixture
def oracle():
    import harness
    return harness.get_backend("oracle")
