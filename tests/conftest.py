import os
import sys
import pytest

os.environ.setdefault("OATK_PF_MIN", "1")     # the host layer's parallel loops use threads even on the small test inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from pyoracle import Ref, have_ref, build
    build()
    if not have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def gpu_ctx():
    from oatk_b200 import lib
    return lib.Context(0)
