"""Stand-ins for the ``chainer.functions`` activations the reference passes to its model specs
(predict_folds.py:144-152 / evaluate.py:96-104: F.sigmoid, F.tanh, F.relu).  They are tags, not
implementations: the arithmetic happens in the GEMM epilogue (csrc/gemm.cu)."""
from __future__ import annotations


class Activation:
    def __init__(self, name):
        self.name = name
        self.__name__ = name

    def __repr__(self):
        return f"F.{self.name}"

    def __call__(self, x):  # pragma: no cover - never evaluated on the host
        raise RuntimeError("activation tags are fused into the device kernels; they are not callable on the host")


relu = Activation("relu")
sigmoid = Activation("sigmoid")
tanh = Activation("tanh")
identity = Activation("identity")

_BY_NAME = {"relu": relu, "sigmoid": sigmoid, "tanh": tanh, "identity": identity}


def resolve(act):
    """Accept an Activation tag, a name, or any callable whose __name__ is relu/sigmoid/tanh."""
    if isinstance(act, Activation):
        return act
    name = act if isinstance(act, str) else getattr(act, "__name__", None)
    if name in _BY_NAME:
        return _BY_NAME[name]
    raise ValueError(f"unsupported activation {act!r}; expected relu, sigmoid or tanh")
