"""pytest configuration: marker registration, import paths, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ann-search-rs_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import annb200
        return annb200.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must never pass on a silent fallback: no device -> hard failure, not a skip."""
    import annb200
    n = annb200.device_count()
    assert n > 0, "no CUDA device visible"
    return n


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(1234)
