"""chainer.variable: ``Variable`` is an ndarray view whose ``.data`` is the plain array, so the reference's
``model(chainer.Variable(x)).data`` (predict_folds.py:55, evaluateModelForTest.py:73) and its operator arithmetic
(``z += U_z(h)``, ``r * h``, ``h /= 2 * K`` -- MGRU.py:71-74, evaluate.py:38-48) run as the fp32 NumPy operations
Chainer's CPU backend performs."""
import numpy as np


class Variable(np.ndarray):
    def __new__(cls, data=None, **_kw):
        if isinstance(data, Parameter):
            data = data.data
        return np.asarray(data).view(cls)

    @property
    def data(self):
        return self.view(np.ndarray)

    array = data

    def to_cpu(self):
        return self

    def to_gpu(self, device=None):
        return self


class Parameter:
    """A named parameter: ``.data`` is an ndarray, or None until the owning link learns its input size (the reference
    builds every link with ``in_size=None``) or ``load_npz`` fills it."""

    def __init__(self, initializer=None, shape=None, name=None):
        self.name = name
        self.initializer = initializer
        self.data = None
        if shape is not None:
            self.initialize(shape)

    def initialize(self, shape):
        init = self.initializer
        if init is None or isinstance(init, (int, float)):
            self.data = np.full(shape, 0.0 if init is None else init, dtype=np.float32)
        elif isinstance(init, np.ndarray):
            self.data = np.array(init, dtype=np.float32).reshape(shape)
        else:
            self.data = np.empty(shape, dtype=np.float32)
            init(self.data)

    @property
    def array(self):
        return self.data

    @property
    def shape(self):
        return None if self.data is None else self.data.shape

    def to_cpu(self):
        return self

    def to_gpu(self, device=None):
        return self
