import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads libvfnerf_b200.so; nvcc cross-compiles without a GPU."""
    from vfnerf_b200 import _lib
    _lib.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def debug_lib():
    """Builds (if stale) and loads the TEST-ONLY libvfnerf_b200_debug.so (UMMA probes, stash read-back)."""
    from vfnerf_b200 import _lib
    _lib.build_debug()
    return _lib.debug_lib()
