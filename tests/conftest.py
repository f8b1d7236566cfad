import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.bindings import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libcpq_ref.so not built (reference tree absent)")
    return Ref()


@pytest.fixture(scope="session")
def checker():
    """Strongest available checker: the compiled reference when present, else the restatement."""
    from oracle.bindings import best_checker
    return best_checker()
