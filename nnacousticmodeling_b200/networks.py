"""Model specs with the reference's constructor / ``__call__`` / ``reset_state`` surface
(scripts/common/chainer_networks.py:8-187) and Chainer ``.npz`` parameter names, computed by the
sm_100a kernels through ``engine``.  Parameters live on the host as NumPy arrays in Chainer layout
(``layer_{l}/W`` (out, in), ``layer_{l}/upward/W`` (4H, in) gate-interleaved, ...); device copies
(bf16 / bf16 hi+lo) are packed lazily per device by ``engine``.
"""
from __future__ import annotations

import sys

import numpy as np

from . import functions as F
from ._native import NnamError

_LSTM_KINDS = ("lstm", "zoneoutlstm", "zoneoutdropoutlstm", "peepholelstm")
_GRU_KINDS = {"gru": (True, "tanh"), "mgrurelu": (False, "relu"), "mgrurelur": (True, "relu")}

_default_precision = "fp32"


def set_default_precision(mode):
    """'fp32' = bf16x3 error-compensated tensor-core GEMMs (<= 1e-3 of the fp32 reference);
    'fp16' / 'bf16' = single-pass 16-bit GEMMs (throughput modes; see engine.Precision for the full grammar)."""
    global _default_precision
    from .engine import Precision
    Precision(mode)  # raises NnamError on an unknown mode
    _default_precision = mode


def _lecun(rng, out_dim, in_dim):
    return (rng.standard_normal((out_dim, in_dim)) / np.sqrt(in_dim)).astype(np.float32)


class Network:
    """Base of all model specs: a flat dict of Chainer-named parameters plus device plans."""

    network = "?"
    recurrent = False
    bidirectional = False

    def __init__(self):
        self.params = {}
        self.precision = _default_precision
        self._version = 0
        self._plans = {}
        self._state = None
        self._device = 0

    # ---- parameter handling -------------------------------------------------------------
    def param_shapes(self, in_size):
        raise NotImplementedError

    def init_params(self, in_size, rng=None):
        """Chainer default initialisers (LeCunNormal weights, zero biases, LSTM forget bias 1)."""
        rng = np.random.default_rng(0) if rng is None else rng
        p = {}
        for name, shape in self.param_shapes(in_size).items():
            if name.endswith("/W") and len(shape) == 2:
                p[name] = _lecun(rng, shape[0], shape[1])
            elif name.endswith("/W"):
                fan_in = int(np.prod(shape[1:]))
                p[name] = (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
            else:
                p[name] = np.zeros(shape, dtype=np.float32)
        if self.network == "lstm":
            for name in p:
                if name.endswith("upward/b"):
                    p[name][2::4] = 1.0  # forget gate rows 4j+2
        self.load_params(p)
        return self

    def load_params(self, params):
        clean = {}
        for k, v in params.items():
            k = k[len("predictor/"):] if k.startswith("predictor/") else k
            clean[k.lstrip("/")] = np.ascontiguousarray(np.asarray(v, dtype=np.float32))
        self.params = clean
        self._validate()
        self._version += 1
        self._plans.clear()
        self.reset_state()

    def _validate(self):
        in_size = self.in_size
        if in_size is None:
            raise NnamError(f"{type(self).__name__}: parameters do not define the input size")
        want = self.param_shapes(in_size)
        for name, shape in want.items():
            if name not in self.params:
                raise NnamError(f"{type(self).__name__}: missing parameter '{name}'")
            if tuple(self.params[name].shape) != tuple(shape):
                raise NnamError(f"{type(self).__name__}: parameter '{name}' has shape "
                                f"{self.params[name].shape}, expected {tuple(shape)}")

    @property
    def in_size(self):
        return None

    def namedparams(self):
        return dict(self.params)

    # ---- Chainer-style device / state surface -----------------------------------------
    def to_gpu(self, device=None):
        self._device = 0 if device is None else int(device)
        return self

    def to_cpu(self):
        raise NnamError("nnacousticmodeling_b200 has no CPU path (gpu < 0 is not supported)")

    def reset_state(self):
        self._state = None

    def __call__(self, x):
        from . import engine
        return engine.call_model(self, x)


class MLP(Network):
    """chainer_networks.py:8-22."""

    network = "ff"

    def __init__(self, n_units, n_out, layers=2, dropout=0, activation=F.relu):
        super().__init__()
        self.n_units, self.n_out, self.layers, self.dropout = int(n_units), int(n_out), int(layers), dropout
        self.activation = F.resolve(activation)

    def param_shapes(self, in_size):
        s, d = {}, in_size
        for l in range(self.layers):
            s[f"layer_{l}/W"] = (self.n_units, d)
            s[f"layer_{l}/b"] = (self.n_units,)
            d = self.n_units
        s["out/W"] = (self.n_out, d)
        s["out/b"] = (self.n_out,)
        return s

    @property
    def in_size(self):
        k = "layer_0/W" if self.layers > 0 else "out/W"
        return self.params[k].shape[1] if k in self.params else None


class TDNN(Network):
    """chainer_networks.py:24-42 (1 x k convolutions along the spliced window; quirk Q3 reshape)."""

    network = "tdnn"

    def __init__(self, n_units, n_out, ksize, dropout=0, activation=F.relu):
        super().__init__()
        if len(n_units) != len(ksize):
            raise ValueError("TDNN n_units argument must have the same length as ksize")
        self.n_units, self.n_out, self.ksize = [int(u) for u in n_units], int(n_out), [int(k) for k in ksize]
        self.layers = len(self.n_units)
        self.dropout = dropout
        self.activation = F.resolve(activation)
        self.input_win_size = sum(self.ksize) - len(self.ksize) + 1

    def param_shapes(self, in_size):
        s, c = {}, in_size // self.input_win_size
        for l, (u, k) in enumerate(zip(self.n_units, self.ksize)):
            s[f"layer_{l}/W"] = (u, c, 1, k)
            s[f"layer_{l}/b"] = (u,)
            c = u
        s["out/W"] = (self.n_out, c)
        s["out/b"] = (self.n_out,)
        return s

    @property
    def in_size(self):
        return self.params["layer_0/W"].shape[1] * self.input_win_size if "layer_0/W" in self.params else None


class _Recurrent(Network):
    recurrent = True

    def __init__(self, n_units, n_out, layers=2, dropout=0, bidirectional=False):
        super().__init__()
        self.n_units, self.n_out, self.layers, self.dropout = int(n_units), int(n_out), int(layers), dropout
        self.bidirectional = bool(bidirectional)

    def _dirs(self):
        return ("fwd/", "bwd/") if self.bidirectional else ("",)

    def _cell_shapes(self, pre, d):
        raise NotImplementedError

    def param_shapes(self, in_size):
        s, d = {}, in_size
        for l in range(self.layers):
            for dd in self._dirs():
                s.update(self._cell_shapes(f"layer_{l}/{dd}", d))
            d = self.n_units * (2 if self.bidirectional else 1)
        s["out/W"] = (self.n_out, d)
        s["out/b"] = (self.n_out,)
        return s

    def _first_key(self):
        raise NotImplementedError

    @property
    def in_size(self):
        k = "layer_0/" + self._dirs()[0] + self._first_key()
        return self.params[k].shape[1] if k in self.params else None


class LSTM(_Recurrent):
    """chainer_networks.py:44-62 (L.LSTM cells).  ``bidirectional=True`` is this build's extension for
    BASELINE config 4 (no reference class; parameters ``layer_{l}/fwd/...`` and ``layer_{l}/bwd/...``)."""

    network = "lstm"
    peephole = False

    def _cell_shapes(self, pre, d):
        h = self.n_units
        s = {pre + "upward/W": (4 * h, d), pre + "upward/b": (4 * h,), pre + "lateral/W": (4 * h, h)}
        if self.peephole:
            for g in ("peep_i", "peep_f", "peep_o"):
                s[pre + g + "/W"] = (h, h)
        return s

    def _first_key(self):
        return "upward/W"


class ZoneoutLSTM(LSTM):
    """chainer_networks.py:64-81; zoneout is the identity at inference, so this is LSTM arithmetic with
    plain-Linear initialisation."""

    network = "zoneoutlstm"

    def __init__(self, n_units, n_out, layers=2, c_ratio=0.5, h_ratio=0.5):
        super().__init__(n_units, n_out, layers, 0)
        self.c_ratio, self.h_ratio = c_ratio, h_ratio


class ZoneoutDropoutLSTM(LSTM):
    """chainer_networks.py:83-101."""

    network = "zoneoutdropoutlstm"

    def __init__(self, n_units, n_out, layers=2, dropout=0, c_ratio=0.5, h_ratio=0.5):
        super().__init__(n_units, n_out, layers, dropout)
        self.c_ratio, self.h_ratio = c_ratio, h_ratio


class PeepholeLSTM(LSTM):
    """chainer_networks.py:103-121 (L.StatefulPeepholeLSTM, full-matrix peepholes)."""

    network = "peepholelstm"
    peephole = True


class NetMGRU(_Recurrent):
    """chainer_networks.py:143-161 with scripts/common/MGRU.py:10-85."""

    network = "mgrurelu"

    def __init__(self, n_units, n_out, layers=2, dropout=0, use_reset_gate=False, activation=F.relu,
                 bidirectional=False):
        super().__init__(n_units, n_out, layers, dropout, bidirectional)
        self.use_reset_gate = bool(use_reset_gate)
        self.activation = F.resolve(activation)
        self.network = {(False, "relu"): "mgrurelu", (True, "relu"): "mgrurelur", (True, "tanh"): "gru"}.get(
            (self.use_reset_gate, self.activation.name), "mgru")

    def _cell_shapes(self, pre, d):
        h = self.n_units
        names = ["W_z", "U_z", "W", "U"] + (["W_r", "U_r"] if self.use_reset_gate else [])
        s = {}
        for n in names:
            s[pre + n + "/W"] = (h, d if n.startswith("W") else h)
            s[pre + n + "/b"] = (h,)
        return s

    def _first_key(self):
        return "W_z/W"


class GRU(NetMGRU):
    """chainer_networks.py:123-141 (L.GRU = StatefulGRU = MGRU with reset gate and tanh)."""

    def __init__(self, n_units, n_out, layers=2, dropout=0, bidirectional=False):
        super().__init__(n_units, n_out, layers, dropout, True, F.tanh, bidirectional)
        self.network = "gru"


def get_nn(network, layers, units, num_classes, activation, tdnn_ksize, dropout=[0]):
    """chainer_networks.py:163-184 (same dispatch, same error behaviour: print + exit(1))."""
    if network == "ff":
        return MLP(units[0], num_classes, layers, dropout[0], activation)
    elif network == "tdnn":
        return TDNN(units, num_classes, tdnn_ksize, dropout[0], activation)
    elif network == "lstm":
        return LSTM(units[0], num_classes, layers, dropout[0])
    elif network == "zoneoutlstm":
        return ZoneoutLSTM(units[0], num_classes, layers, *dropout)
    elif network == "zoneoutdropoutlstm":
        return ZoneoutDropoutLSTM(units[0], num_classes, layers, *dropout)
    elif network == "peepholelstm":
        return PeepholeLSTM(units[0], num_classes, layers, dropout[0])
    elif network == "gru":
        return GRU(units[0], num_classes, layers, dropout[0])
    elif network == "mgrurelu":
        return NetMGRU(units[0], num_classes, layers, dropout[0], False, F.relu)
    elif network == "mgrurelur":
        return NetMGRU(units[0], num_classes, layers, dropout[0], True, F.relu)
    # extensions of this build for BASELINE config 4 (no reference class)
    elif network == "blstm":
        return LSTM(units[0], num_classes, layers, dropout[0], bidirectional=True)
    elif network == "bgru":
        return GRU(units[0], num_classes, layers, dropout[0], bidirectional=True)
    else:
        print("Wrong network type specified")
        sys.exit(1)


def is_nn_recurrent(n):
    """chainer_networks.py:186-187, plus this build's bidirectional tags."""
    return n.endswith("lstm") or n.startswith("gru") or n.startswith("mgru") or n in ("blstm", "bgru")


class RPL4:
    """scripts/common/RPL.py:58-74: parameters W, b (init 0) and lb (init -20), each (1, C)."""

    def __init__(self, n_out):
        self.n_out = int(n_out)
        self.params = {"W": np.zeros((1, n_out), np.float32), "b": np.zeros((1, n_out), np.float32),
                       "lb": np.full((1, n_out), -20.0, np.float32)}

    def load_params(self, params):
        """Strict, like ``serializers.load_npz`` (a missing parameter is an error there too): an archive that lacks
        W, b or lb would otherwise silently evaluate with the inert initial values."""
        got = {(k[len("predictor/"):] if k.startswith("predictor/") else k): v for k, v in params.items()}
        missing = [k for k in self.params if k not in got]
        if missing:
            raise NnamError(f"RPL4: parameter(s) {missing} missing from the archive (has {sorted(got)})")
        for k in self.params:
            v = np.asarray(got[k], np.float32)
            if v.size != self.n_out:
                raise NnamError(f"RPL4: parameter {k} has {v.size} entries, expected {self.n_out}")
            self.params[k] = v.reshape(1, self.n_out)

    def namedparams(self):
        return dict(self.params)


class Classifier:
    """Shape of ``L.Classifier(model)`` as used by the reference only to carry the ``predictor/`` key
    prefix through ``serializers.load_npz`` (predict_folds.py:157,206; train.py:393-395)."""

    def __init__(self, predictor):
        self.predictor = predictor

    def to_gpu(self, device=None):
        if hasattr(self.predictor, "to_gpu"):
            self.predictor.to_gpu(device)
        return self


def load_npz(path, obj):
    """chainer.serializers.load_npz for the key layout of SURVEY 8b."""
    target = obj.predictor if isinstance(obj, Classifier) else obj
    with np.load(str(path)) as z:
        params = {k: z[k] for k in z.files}
    if isinstance(obj, Classifier):
        params = {k[len("predictor/"):]: v for k, v in params.items() if k.startswith("predictor/")}
    target.load_params(params)
    return obj


def save_npz(path, obj):
    """chainer.serializers.save_npz: ``predictor/``-prefixed keys when given a Classifier."""
    target = obj.predictor if isinstance(obj, Classifier) else obj
    prefix = "predictor/" if isinstance(obj, Classifier) else ""
    np.savez(str(path), **{prefix + k: v for k, v in target.namedparams().items()})
