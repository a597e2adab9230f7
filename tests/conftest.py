import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: long-running statistical test")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Builds the test oracle, the host simulator and (when the reference tree is mounted) oracle/_ref."""
    import __graft_entry__ as g
    g.build_test_infrastructure()
    yield
