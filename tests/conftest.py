import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DECKS = os.path.join(ROOT, "tests", "golden", "decks")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def decks_dir():
    return DECKS


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import build_oracle
    return build_oracle()
