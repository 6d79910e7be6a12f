"""chainer.training: import-time stand-ins (predict_folds.py:14-15 and train.py import these; the hot path never calls
them).  Anything that would actually train raises."""
from . import extension, extensions, trigger, triggers, util  # noqa: F401


def make_extension(trigger=None, priority=None, **_kw):
    def deco(fn):
        fn.trigger, fn.priority = trigger, priority
        return fn
    return deco


class StandardUpdater:
    def __init__(self, *_a, **_k):
        raise NotImplementedError("training is out of scope of the chainer shim")


class Trainer(StandardUpdater):
    pass
