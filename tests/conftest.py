import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope='session')
def cuda_lib():
    """The product library; (re)built in-tree when stale.  Never falls back to anything else."""
    from nerfstyle_b200 import build as nb
    nb.build()
    from nerfstyle_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope='session')
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
