"""NumPy-only stand-in for the parts of Chainer 3.5 the reference's hot path touches (tests/chainer_shim/README.md).

TEST INFRASTRUCTURE ONLY: lets /root/reference/scripts/{common,util}/*.py be imported and run UNMODIFIED on the CPU so
that oracle/make_golden_nets.py can generate golden outputs from the reference's own control flow.  Forward
(inference) semantics only; there is no autograd, no optimiser, no trainer.
"""
import contextlib

import numpy as _np

from . import variable  # noqa: F401
from .variable import Parameter, Variable  # noqa: F401
from . import link  # noqa: F401
from .link import Chain, ChainList, Link  # noqa: F401
from . import cuda, initializers, serializers, reporter, dataset, datasets, iterators, optimizers  # noqa: F401
from . import functions, links, training  # noqa: F401

__version__ = "3.5.0-shim"


class _Config:
    """chainer.config: only ``train`` matters on the path (dropout / zoneout are the identity when it is False;
    predict_folds.py:141, evaluate.py:94, evaluateModelForTest.py:46)."""
    train = True
    enable_backprop = True


config = _Config()
global_config = config


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)


def no_backprop_mode():
    """predict_folds.py:54,85: there is no graph to switch off here."""
    return using_config("enable_backprop", False)


def as_array(x):
    return x.data if isinstance(x, (Variable, Parameter)) else _np.asarray(x)
