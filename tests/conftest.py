import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def product():
    import ballbox_b200
    return ballbox_b200.load()


@pytest.fixture(scope="session")
def checkers():
    """CPU checkers: the compiled reference (if built) and our C restatement."""
    from tests import oracles
    return oracles.load_all()


@pytest.fixture(scope="session")
def truth(checkers):
    """The strongest available checker: oracle/_ref when present, else the restatement."""
    return checkers.get("ref") or checkers["oracle"]
