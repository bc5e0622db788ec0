import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (test infrastructure)."""
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (skipped when it is not available)."""
    from oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libfsref.so not built and /root/reference not mounted")
    return Reference()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
