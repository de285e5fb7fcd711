import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")
    config.addinivalue_line("markers", "multigpu: needs at least two CUDA devices (skipped on a one-GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
