import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


@pytest.fixture(scope='session', autouse=True)
def _extension_built():
    """The .so files are git-ignored build outputs: on a fresh checkout the suite compiles the
    CUDA extension once (nvcc cross-compiles without a GPU).  Only when MISSING -- an existing
    library is never rebuilt behind the tests' back.  (The product itself still fails loudly
    when the library is absent: _native.lib() raises NativeError.)"""
    from reversible_raytracer_b200 import _native as nat
    if not os.path.exists(nat.LIB_PATH) or not os.path.exists(nat.BENCH_LIB_PATH):
        nat.build(force=True)
    yield
