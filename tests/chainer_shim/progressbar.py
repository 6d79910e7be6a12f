"""Stand-in for the `progressbar` package (predict_folds.py:47-48,63-64; evaluateModelForTest.py:64-65,80-81)."""


class ProgressBar:
    def __init__(self, max_value=None, **_kw):
        self.max_value, self.value = max_value, 0

    def update(self, value=None):
        if value is not None:
            self.value = value

    def __iadd__(self, n):
        self.value += n
        return self

    def finish(self):
        pass
