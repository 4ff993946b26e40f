import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def opz():
    """The product package (ctypes binding of libopz.so + host mirror).  Import failure = test failure."""
    import opz_import
    return opz_import.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import nde_oracle
    return nde_oracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
