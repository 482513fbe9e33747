import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
# the GPU tests must run on the hand-written kernels: any drop to a torch / cuDNN library path raises (kgc_gcn_b200._lib.library_path)
os.environ.setdefault('KGC_STRICT', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def toy_dir():
    return os.path.join(GOLDEN, 'data', 'Toy')
