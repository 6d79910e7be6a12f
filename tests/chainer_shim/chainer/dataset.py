"""chainer.dataset: base class for orcus_chainer_util.SequenceShuffleIterator (training only)."""


class Iterator:
    pass


def concat_examples(batch, device=None, padding=None):
    raise NotImplementedError("training is out of scope of the chainer shim")
