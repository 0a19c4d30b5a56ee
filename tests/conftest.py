import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def qps():
    from apf_quadruped_b200 import qpswift
    qpswift.lib()  # fails loudly if the C-ABI library has not been built
    return qpswift


@pytest.fixture(scope="session")
def oracle():
    from oracle import reference
    if not (reference.available("reference") or reference.available("port")):
        pytest.skip("no oracle library built (make -C oracle)")
    return reference
