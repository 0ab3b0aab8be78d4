"""tf.summary no-ops that record what was logged (test infrastructure only)."""
import numpy as np

from ._tensor import _arr

LOG = {}


def scalar(name, tensor):
    LOG[name] = np.asarray(_arr(tensor))


def histogram(name, tensor):
    LOG[name] = np.asarray(_arr(tensor))
