import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def huf():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("huffman-avx512_b200")


@pytest.fixture(scope="session")
def oracle():
    from _libs import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from _libs import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libhufref.so not built (reference sources absent)")
    return Ref()
