"""pytest configuration: markers and shared fixtures."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger sizes; still part of the default runs")


@pytest.fixture
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def lib():
    from pgsd_sph_b200 import _lib
    return _lib.load()


def _cuda_available():
    try:
        from pgsd_sph_b200 import _lib
        return bool(_lib.load().pgsd_b200_cuda_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a box without a GPU: they are deselected by `-m "not gpu"`
    # in the CPU run; if someone runs them here without a device they fail loudly in the test body.
    pass
