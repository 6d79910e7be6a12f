"""chainer.initializers: Constant (RPL.py:61-66) and the LeCunNormal default of L.Linear
(chainer/initializers/normal.py: std = scale / sqrt(fan_in))."""
import numpy as np


class Constant:
    def __init__(self, fill_value):
        self.fill_value = fill_value

    def __call__(self, array):
        array[...] = self.fill_value


class LeCunNormal:
    def __init__(self, scale=1.0, rng=None):
        self.scale, self.rng = scale, rng

    def __call__(self, array):
        fan_in = int(np.prod(array.shape[1:]))
        rng = self.rng or np.random
        array[...] = (rng.standard_normal(array.shape) * (self.scale / np.sqrt(fan_in))).astype(array.dtype)


Zero = lambda: Constant(0.0)  # noqa: E731
One = lambda: Constant(1.0)  # noqa: E731
