import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


@pytest.fixture(scope='session', autouse=True)
def native_libraries():
    """Build the in-tree libraries (nvcc cross-compiles without a GPU) if a fresh checkout lacks them."""
    pkg = os.path.join(ROOT, 'ballermixplus_b200')
    needed = [os.path.join(pkg, 'libblmx.so'), os.path.join(pkg, 'libblmx_checked.so'),
              os.path.join(pkg, 'libblmx_io.so'), os.path.join(ROOT, 'oracle', 'liboracle.so')]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
