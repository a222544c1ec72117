import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def sme():
    """The product package; the in-tree library is (re)built first so a stale .so cannot be tested."""
    import __graft_entry__ as ge
    ge._load_build_module().build()
    import scimlexpectations_jl_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def gpu(sme):
    if sme.device_count() < 1:
        pytest.fail("marked gpu but no CUDA device is visible (there is no CPU fallback to fall back on)")
    return sme
