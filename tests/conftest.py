import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """The CPU oracle (oracle/libnd4ref.so) — the checker, never the thing under test."""
    from oracle import nd4ref
    nd4ref.build()
    return nd4ref


@pytest.fixture(scope="session")
def la():
    """The product: nd4js_b200.la on the GPU. Fails loudly when no device / library is present."""
    import nd4js_b200
    nd4js_b200.init()
    return nd4js_b200.la
