import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a hard limit (pytest-timeout, thread method: the process exits even when the host thread is
    blocked inside a CUDA call), so that a kernel that never finishes fails the run instead of holding the box."""
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(600, method="thread"))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _pin_label_path(monkeypatch):
    """The NIW label path adapts at run time (tensor-core kernel <-> FMA kernel, from the previous call's candidate
    counters).  Parity tests pin it so that each test exercises the path it names; the adaptive switch has its own
    test (test_label_path_adapts_to_overlapping_clusters)."""
    monkeypatch.setenv("DPMM_LABEL_ADAPT", "0")
