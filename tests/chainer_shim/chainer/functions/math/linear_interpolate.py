"""chainer/functions/math/linear_interpolate.py (CPU forward): ``one = p.dtype.type(1); y = p * x + (one - p) * y``."""
import numpy as np

from ...variable import Variable


def linear_interpolate(p, x, y):
    p, x, y = (np.asarray(t.data if isinstance(t, Variable) else t) for t in (p, x, y))
    one = p.dtype.type(1)
    return (p * x + (one - p) * y).view(Variable)
