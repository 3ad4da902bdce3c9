import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("hevc-deep-learning-pipeline_b200")


@pytest.fixture(scope="session")
def host():
    return importlib.import_module("hevc-deep-learning-pipeline_b200.host")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def weights(oracle, host):
    return oracle.load_weights(host.DEFAULT_WEIGHTS)


@pytest.fixture(scope="session")
def built():
    """Make sure libhevcdl.so exists (GPU tests call through the C-ABI)."""
    import __graft_entry__ as ge
    so = os.path.join(ge.CSRC, "libhevcdl.so")
    if not os.path.exists(so):
        ge.build()
    return so
