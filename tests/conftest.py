import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def g13():
    return dict(np.load(os.path.join(GOLDEN, 'g1_g3_bivariate_gaussian.npz')))


@pytest.fixture(scope='session')
def g45():
    return dict(np.load(os.path.join(GOLDEN, 'g4_g5_gaussian_mixture.npz')))


@pytest.fixture(scope='session')
def ref_test_ksd():
    return dict(np.load(os.path.join(GOLDEN, 'reference_test_ksd.npz')))


@pytest.fixture(scope='session')
def gpu_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return 'cuda:0'
