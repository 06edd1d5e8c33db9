import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TOOLS = os.path.join(ROOT, "tests", "tools")   # test-only helpers (hostemu: host emulation of the backward kernels)
if TOOLS not in sys.path:
    sys.path.insert(0, TOOLS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from crfp_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from crfp_b200.build import build
        build()
    return _lib.lib()
