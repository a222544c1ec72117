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


# ---- observed parity errors: every GPU parity test may record what it measured next to what it asserts; the
# session writes them to gpurun_out/parity_observed.json (copied into profiles/ per round) ----
_OBSERVED = {}


def record_parity(test: str, **values):
    """Remember the worst observed value of each quantity for `test` (floats: maximum; everything else: last)."""
    slot = _OBSERVED.setdefault(test, {})
    for k, v in values.items():
        if isinstance(v, (int, float)) and not isinstance(v, bool) and isinstance(slot.get(k), (int, float)):
            slot[k] = max(slot[k], v)
        else:
            slot[k] = v


def pytest_sessionfinish(session, exitstatus):
    if not _OBSERVED:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_observed.json"), "w") as f:
            json.dump({"exitstatus": int(exitstatus), "observed": _OBSERVED}, f, indent=1, sort_keys=True, default=float)
    except OSError:
        pass
