"""chainer.datasets (train.py:267-269; training only)."""


class TupleDataset:
    def __init__(self, *arrays):
        self.arrays = arrays

    def __len__(self):
        return len(self.arrays[0])

    def __getitem__(self, i):
        return tuple(a[i] for a in self.arrays)
