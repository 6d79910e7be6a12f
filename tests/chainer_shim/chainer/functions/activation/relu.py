"""chainer/functions/activation/relu.py (CPU forward): ``numpy.maximum(x, 0, dtype=x.dtype)``."""
import numpy as np

from ...variable import Variable


def relu(x):
    x = x.data if isinstance(x, Variable) else np.asarray(x)
    return np.maximum(x, 0, dtype=x.dtype).view(Variable)
