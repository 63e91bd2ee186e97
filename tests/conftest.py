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


def pytest_collection_modifyitems(config, items):
    """A GPU run on a tree without built artefacts (a plain git checkout on the box; `*.so` is git-ignored):
    build the CUDA library once up front -- nvcc is in the image.  With the artefacts present (the normal
    case: they travel with the snapshot) this does nothing; on the CPU it does nothing either, the product
    path keeps failing loudly when the library is missing (tests/test_abi.py)."""
    if not any(it.get_closest_marker("gpu") for it in items):
        return
    from cpfft_b200.api import library_path
    if not os.path.exists(library_path()):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            from cpfft_b200.build import build
            build()
