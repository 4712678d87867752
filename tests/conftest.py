import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _pin_label_path(monkeypatch):
    """The NIW label path adapts at run time (tensor-core kernel <-> FMA kernel, from the previous call's candidate
    counters).  Parity tests pin it so that each test exercises the path it names; the adaptive switch has its own
    test (test_label_path_adapts_to_overlapping_clusters)."""
    monkeypatch.setenv("DPMM_LABEL_ADAPT", "0")
