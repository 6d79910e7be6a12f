"""chainer/functions/activation/sigmoid.py (CPU forward): ``half = x.dtype.type(0.5);
y = tanh(x * half) * half + half``."""
import numpy as np

from ...variable import Variable


def sigmoid(x):
    x = x.data if isinstance(x, Variable) else np.asarray(x)
    half = x.dtype.type(0.5)
    return (np.tanh(x * half) * half + half).view(Variable)
