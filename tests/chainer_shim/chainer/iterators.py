"""chainer.iterators (imported by predict_folds.py:16 / train.py:15; training only)."""


class SerialIterator:
    def __init__(self, *_a, **_k):
        raise NotImplementedError("training is out of scope of the chainer shim")
