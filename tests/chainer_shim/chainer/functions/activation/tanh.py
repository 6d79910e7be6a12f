"""chainer/functions/activation/tanh.py (CPU forward): ``numpy.tanh(x)``."""
import numpy as np

from ...variable import Variable


def tanh(x):
    x = x.data if isinstance(x, Variable) else np.asarray(x)
    return np.tanh(x).view(Variable)
