import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
PKG = 'single-shot-detector_b200'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_pkg(sub=None):
    """The package directory name has a hyphen, so it is imported by string."""
    return importlib.import_module(PKG if sub is None else PKG + '.' + sub)


@pytest.fixture(scope='session')
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name + '.npz'))
    return _load


@pytest.fixture(scope='session')
def syn():
    return load_pkg('synthetic')
