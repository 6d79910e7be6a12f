"""chainer.optimizers (train.py:297-301; training only)."""


def _no_training(*_a, **_k):
    raise NotImplementedError("training is out of scope of the chainer shim")


SGD = MomentumSGD = Adam = _no_training
