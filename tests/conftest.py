import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box via gpurun)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """The CUDA library is built in-tree (git-ignored): (re)build it when missing or stale.
    nvcc cross-compiles without a GPU, so this also works in the build container."""
    try:
        from gym_copter_b200 import build
        build.build()
    except Exception as e:      # leave it to the tests that need it to fail loudly
        print('libcopter_b200.so build skipped: %r' % (e,))
    yield
