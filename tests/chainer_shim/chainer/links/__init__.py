"""chainer.links used by the reference model specs (chainer_networks.py).  Each ``__call__`` restates the forward of
the Chainer v3.5 link it names (upstream path cited; recalled, not vendored) in a few NumPy lines."""
import numpy as np

from .. import functions as F
from ..link import Chain, Link
from ..variable import Variable
from . import connection  # noqa: F401
from .connection.linear import Linear  # noqa: F401


def _zeros_like_rows(x, size):
    return np.zeros((len(x), size), dtype=np.asarray(x).dtype).view(Variable)


class LSTM(Chain):
    """chainer/links/connection/lstm.py (class LSTM): ``upward`` Linear(in, 4H) with bias, ``lateral`` Linear(H, 4H,
    nobias); ``lstm_in = upward(x); if h is not None: lstm_in += lateral(h); if c is None: c = zeros;
    c, h = F.lstm(c, lstm_in)``.  ``L.LSTM(None, n)``: input size learnt at the first call."""

    def __init__(self, in_size, out_size=None, **_kw):
        super().__init__()
        if out_size is None:
            in_size, out_size = None, in_size
        self.state_size = out_size
        with self.init_scope():
            self.upward = Linear(in_size, 4 * out_size)
            self.lateral = Linear(out_size, 4 * out_size, nobias=True)
        self.reset_state()

    def reset_state(self):
        self.c = self.h = None

    def __call__(self, x):
        lstm_in = self.upward(x)
        if self.h is not None:
            lstm_in += self.lateral(self.h)
        if self.c is None:
            self.c = _zeros_like_rows(x, self.state_size)
        self.c, self.h = F.lstm(self.c, lstm_in)
        return self.h


class StatefulZoneoutLSTM(Chain):
    """chainer/links/connection/zoneoutlstm.py: same parameters as LSTM (plain Linear initialisers); ``lstm_in =
    upward(x); if h is not None: lstm_in += lateral(h) else: h = zeros; if c is None: c = zeros;
    c_next, h_next = F.lstm(c, lstm_in); c = F.zoneout(c, c_next, c_ratio); h = F.zoneout(h, h_next, h_ratio)`` --
    at inference zoneout returns the new value."""

    def __init__(self, in_size, out_size, c_ratio=0.5, h_ratio=0.5, **_kw):
        super().__init__()
        self.state_size, self.c_ratio, self.h_ratio = out_size, c_ratio, h_ratio
        with self.init_scope():
            self.upward = Linear(in_size, 4 * out_size)
            self.lateral = Linear(out_size, 4 * out_size, nobias=True)
        self.reset_state()

    def reset_state(self):
        self.c = self.h = None

    def __call__(self, x):
        lstm_in = self.upward(x)
        if self.h is not None:
            lstm_in += self.lateral(self.h)
        else:
            self.h = _zeros_like_rows(x, self.state_size)
        if self.c is None:
            self.c = _zeros_like_rows(x, self.state_size)
        c_next, h_next = F.lstm(self.c, lstm_in)
        self.c = F.zoneout(self.c, c_next, self.c_ratio)
        self.h = F.zoneout(self.h, h_next, self.h_ratio)
        return self.h


class StatefulPeepholeLSTM(Chain):
    """chainer/links/connection/peephole.py: ``upward`` Linear(in, 4H), ``lateral`` Linear(H, 4H, nobias), ``peep_i``,
    ``peep_f``, ``peep_o`` Linear(H, H, nobias) -- FULL matrices.  Gates are split as ``reshape(B, H, 4)[:, :, k]``
    (a, i, f, o); ``a = tanh(a); i = sigmoid(i + peep_i(c)); f = sigmoid(f + peep_f(c)); c_next = a * i + f * c;
    o = sigmoid(o + peep_o(c_next)); h = o * tanh(c_next)``."""

    def __init__(self, in_size, out_size):
        super().__init__()
        self.state_size = out_size
        with self.init_scope():
            self.upward = Linear(in_size, 4 * out_size)
            self.lateral = Linear(out_size, 4 * out_size, nobias=True)
            self.peep_i = Linear(out_size, out_size, nobias=True)
            self.peep_f = Linear(out_size, out_size, nobias=True)
            self.peep_o = Linear(out_size, out_size, nobias=True)
        self.reset_state()

    def reset_state(self):
        self.c = self.h = None

    def __call__(self, x):
        lstm_in = self.upward(x)
        if self.h is not None:
            lstm_in += self.lateral(self.h)
        if self.c is None:
            self.c = _zeros_like_rows(x, self.state_size)
        g = lstm_in.data.reshape(len(lstm_in), lstm_in.shape[1] // 4, 4)
        a, i, f, o = (g[:, :, k].view(Variable) for k in range(4))
        a = F.tanh(a)
        i = F.sigmoid(i + self.peep_i(self.c))
        f = F.sigmoid(f + self.peep_f(self.c))
        c_next = a * i + f * self.c
        o = F.sigmoid(o + self.peep_o(c_next))
        self.c = c_next
        self.h = o * F.tanh(c_next)
        return self.h


class StatefulGRU(Chain):
    """chainer/links/connection/gru.py (class StatefulGRU; ``L.GRU`` in v3): six Linears ``W_r, U_r, W_z, U_z, W, U``,
    all with bias.  ``z = W_z(x); h_bar = W(x); if h is not None: r = sigmoid(W_r(x) + U_r(h)); z += U_z(h);
    h_bar += U(r * h); z = sigmoid(z); h_bar = tanh(h_bar); h_new = linear_interpolate(z, h_bar, h) if h is not None
    else z * h_bar``.  (The reference's MGRU.py:67-85 is this step with two options added.)"""

    def __init__(self, in_size, out_size, **_kw):
        super().__init__()
        self.state_size = out_size
        with self.init_scope():
            self.W_r = Linear(in_size, out_size)
            self.U_r = Linear(out_size, out_size)
            self.W_z = Linear(in_size, out_size)
            self.U_z = Linear(out_size, out_size)
            self.W = Linear(in_size, out_size)
            self.U = Linear(out_size, out_size)
        self.reset_state()

    def reset_state(self):
        self.h = None

    def __call__(self, x):
        z = self.W_z(x)
        h_bar = self.W(x)
        if self.h is not None:
            r = F.sigmoid(self.W_r(x) + self.U_r(self.h))
            z += self.U_z(self.h)
            h_bar += self.U(r * self.h)
        z = F.sigmoid(z)
        h_bar = F.tanh(h_bar)
        self.h = F.linear_interpolate(z, h_bar, self.h) if self.h is not None else z * h_bar
        return self.h


GRU = StatefulGRU


class Convolution2D(Link):
    """chainer/links/connection/convolution_2d.py: W (out, in, kh, kw) LeCunNormal, b (out,) zeros; in_channels=None is
    learnt at the first call; forward = F.convolution_2d (cross-correlation)."""

    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False, **_kw):
        super().__init__()
        from ..initializers import LeCunNormal
        from ..variable import Parameter
        self.ksize = (ksize, ksize) if isinstance(ksize, int) else tuple(ksize)
        self.stride, self.pad, self.out_channels = stride, pad, out_channels
        with self.init_scope():
            self.W = Parameter(LeCunNormal())
            self.b = None if nobias else Parameter(0.0, (out_channels,))
        if in_channels is not None:
            self.W.initialize((out_channels, in_channels) + self.ksize)

    def __call__(self, x):
        if self.W.data is None:
            self.W.initialize((self.out_channels, x.shape[1]) + self.ksize)
        return F.convolution_2d(x, self.W, self.b, self.stride, self.pad)


class Classifier(Chain):
    """chainer/links/model/classifier.py: wraps the model as child ``predictor`` -- the reason every key of the
    reference's .npz files starts with ``predictor/`` (train.py:393-395).  Loss / accuracy are training-only."""

    def __init__(self, predictor, lossfun=None, accfun=None):
        super().__init__()
        with self.init_scope():
            self.predictor = predictor

    def __call__(self, *args):
        raise NotImplementedError("training losses are out of scope of the chainer shim")
