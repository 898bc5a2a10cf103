import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import harness
    return harness.oracle_lib()


@pytest.fixture(scope="session")
def cuda():
    import harness
    return harness.cuda_lib()
