import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def nib():
    """The product package with libnib.so built (nvcc cross-compiles here; no GPU needed to build)."""
    from network_interpretation_imagenet_b200 import build as _b
    _b.build()
    import network_interpretation_imagenet_b200 as pkg
    return pkg
