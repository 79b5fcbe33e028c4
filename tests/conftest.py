import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def lib():
    """Build (if needed) and load libdomainrag_b200.so."""
    from domain_rag_b200 import _lib, build
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def golden_dir():
    return REPO / "tests" / "golden"
