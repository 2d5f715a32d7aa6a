import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    """Fixtures generated FROM THE REFERENCE'S OWN CODE by tests/golden/make_golden.py."""
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.npz'))


@pytest.fixture(scope='session')
def lib():
    from gabotorch_b200 import _lib
    return _lib.load()
