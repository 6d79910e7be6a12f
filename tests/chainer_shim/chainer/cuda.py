"""chainer.cuda on a machine without CuPy: arrays stay NumPy arrays (predict_folds.py:53,56,84,87,158-161)."""
import numpy as np

available = False


class _Device:
    def use(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def get_device_from_id(_id):
    return _Device()


get_device = get_device_from_id


def to_gpu(array, device=None, stream=None):
    return array


def to_cpu(array, stream=None):
    from .variable import Variable
    return array.data if isinstance(array, Variable) else array


def get_array_module(*_args):
    return np
