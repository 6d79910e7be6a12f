"""chainer.functions used on the reference's hot path, as the NumPy expressions Chainer v3.5's CPU ``forward``
methods evaluate (upstream paths cited per function; recalled, not vendored).  All take / return ``Variable``."""
import numpy as np

from ..variable import Parameter, Variable
from . import activation, math  # noqa: F401
from .activation.relu import relu  # noqa: F401
from .activation.sigmoid import sigmoid  # noqa: F401
from .activation.tanh import tanh  # noqa: F401
from .math.linear_interpolate import linear_interpolate  # noqa: F401


def _a(x):
    return x.data if isinstance(x, (Variable, Parameter)) else np.asarray(x)


def _v(x):
    return np.asarray(x).view(Variable)


def linear(x, W, b=None):
    """chainer/functions/connection/linear.py: ``y = x.dot(W.T).astype(x.dtype); y += b``; inputs with more than two
    axes are flattened per sample first."""
    x, W = _a(x), _a(W)
    if x.ndim > 2:
        x = x.reshape(len(x), -1)
    y = x.dot(W.T).astype(x.dtype, copy=False)
    if b is not None:
        y += _a(b)
    return _v(y)


def dropout(x, ratio=.5):
    """chainer/functions/noise/dropout.py: the identity unless ``chainer.config.train``."""
    import chainer
    if chainer.config.train and ratio > 0:
        raise NotImplementedError("training-mode dropout is out of scope of the chainer shim")
    return _v(_a(x))


def zoneout(h, x, ratio=.5):
    """chainer/functions/noise/zoneout.py: ``if configuration.config.train: ...; return x`` -- the NEW value at
    inference."""
    import chainer
    if chainer.config.train:
        raise NotImplementedError("training-mode zoneout is out of scope of the chainer shim")
    return _v(_a(x))


def reshape(x, shape):
    return _v(_a(x).reshape(shape))


def broadcast_to(x, shape):
    return _v(np.broadcast_to(_a(x), shape))


def maximum(a, b):
    return _v(np.maximum(_a(a), _a(b)))


def minimum(a, b):
    return _v(np.minimum(_a(a), _a(b)))


def exp(x):
    return _v(np.exp(_a(x)))


def log(x):
    return _v(np.log(_a(x)))


def log_softmax(x, axis=1):
    """chainer/functions/activation/log_softmax.py (CPU): ``log_z = logsumexp(x)`` with
    ``m = x.max(axis=1, keepdims=True); y = x - m; exp(y, out=y); s = y.sum(axis=1, keepdims=True); log(s, out=s);
    m += s``, then ``y = x - log_z``."""
    x = _a(x)
    m = x.max(axis=axis, keepdims=True)
    y = x - m
    np.exp(y, out=y)
    s = y.sum(axis=axis, keepdims=True)
    np.log(s, out=s)
    m = m + s
    return _v(x - m)


def lstm(c_prev, x):
    """chainer/functions/activation/lstm.py (CPU): ``a, i, f, o = _extract_gates(x)`` where
    ``_extract_gates(x) = x.reshape((len(x), x.shape[1] // 4, 4) + x.shape[2:])[:, :, k]``; ``a = tanh(a)``,
    ``i, f, o = _sigmoid(.)`` with ``_sigmoid(x) = tanh(x * 0.5) * 0.5 + 0.5``; ``c = a * i + f * c_prev``;
    ``h = o * tanh(c)``.  Returns ``(c, h)``."""
    c_prev, x = _a(c_prev), _a(x)
    r = x.reshape((len(x), x.shape[1] // 4, 4) + x.shape[2:])
    half = x.dtype.type(0.5)
    a = np.tanh(r[:, :, 0])
    i = np.tanh(r[:, :, 1] * half) * half + half
    f = np.tanh(r[:, :, 2] * half) * half + half
    o = np.tanh(r[:, :, 3] * half) * half + half
    c = a * i + f * c_prev
    h = o * np.tanh(c)
    return _v(c), _v(h)


def convolution_2d(x, W, b=None, stride=1, pad=0):
    """chainer/functions/connection/convolution_2d.py (CPU): im2col, ``tensordot(col, W, ((1, 2, 3), (1, 2, 3)))``,
    ``+ b``, ``rollaxis(y, 3, 1)`` -- a cross-correlation (no kernel flip).  Unit stride, no padding (all the
    reference's TDNN uses, chainer_networks.py:35)."""
    x, W = _a(x), _a(W)
    if stride not in (1, (1, 1)) or pad not in (0, (0, 0)):
        raise NotImplementedError("chainer shim: convolution_2d with stride 1 / pad 0 only")
    kh, kw = W.shape[2], W.shape[3]
    oh, ow = x.shape[2] - kh + 1, x.shape[3] - kw + 1
    col = np.empty((x.shape[0], x.shape[1], kh, kw, oh, ow), dtype=x.dtype)
    for j in range(kh):
        for i in range(kw):
            col[:, :, j, i] = x[:, :, j:j + oh, i:i + ow]
    y = np.tensordot(col, W, ((1, 2, 3), (1, 2, 3))).astype(x.dtype, copy=False)
    if b is not None:
        y += _a(b)
    return _v(np.rollaxis(y, 3, 1))


def softmax_cross_entropy(*_a_, **_k):
    raise NotImplementedError("training losses are out of scope of the chainer shim")


accuracy = softmax_cross_entropy
