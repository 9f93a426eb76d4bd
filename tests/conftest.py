import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The real reference compiled for this host; skips where oracle/_ref/libsvo_ref.so was never built."""
    from oracle.pyoracle import Ref
    r = Ref()
    if not r.available():
        pytest.skip("oracle/_ref/libsvo_ref.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def ctx():
    """CUDA context of the product library. GPU tests fail (not skip) if the library or device is missing."""
    from android_svo_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
