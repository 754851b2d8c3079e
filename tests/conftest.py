import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    from noir_backend_using_gnark_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def ctx(lib_built):
    import noir_backend_using_gnark_b200 as zk

    c = zk.Context(0)
    yield c
    c.close()
