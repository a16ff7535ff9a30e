import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_api():
    from oracle import orc
    return orc.api()


@pytest.fixture(scope="session")
def gpu_api():
    from dbox_b200 import lib
    a = lib.api()   # raises if the CUDA library is missing: there is no fallback to test instead
    if a.device_count() < 1:
        pytest.fail("no CUDA device visible but a gpu-marked test was selected")
    return a
