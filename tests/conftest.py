import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='session')
def oracle():
    import ssn_oracle
    ssn_oracle.lib()          # builds oracle/_build on first use
    return ssn_oracle


@pytest.fixture(scope='session')
def built_library():
    """The CUDA library, compiled here if the .so is absent (nvcc cross-compiles)."""
    path = os.path.join(ROOT, 'tc_gan_b200', 'ext', 'libssnode.so')
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    return path
