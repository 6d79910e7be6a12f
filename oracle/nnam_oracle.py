"""CPU oracle for the network-output hot path of OrcusCZ/NNAcousticModeling.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nnacousticmodeling_b200/`` may import this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

It is a NumPy fp32 restatement of what the reference computes between "feature matrix"
and "per-frame pdf log-likelihoods".  Citations are ``file:line`` relative to the
reference checkout (``/root/reference``).

Pinning status
--------------
* splice / feature transform / logsum / time delay / .lab writer: PINNED.  The reference's
  own NumPy helpers import and run; ``oracle/make_golden.py`` ran them and committed their
  outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this module
  against those vectors bit for bit.
* network cells, predict(), NNWithRPL, RPL4, evaluateModelTestTri: PINNED (round 2) to
  outputs of the reference's OWN code.  ``oracle/make_golden_nets.py`` runs the unmodified
  ``chainer_networks.py`` / ``MGRU.py`` / ``RPL.py`` / ``predict_folds.py`` / ``evaluate.py``
  / ``evaluateModelForTest.py`` / ``master_script.py`` on a NumPy-only stand-in for the
  few Chainer 3.5 classes they touch (``tests/chainer_shim``; Chainer itself is neither
  vendored nor installable here) and commits inputs + outputs under ``tests/golden/nets/``;
  ``tests/test_reference_goldens.py`` checks this module against them (< 2e-5) for all nine
  ``get_nn`` kinds, fold / dev mode and every folds / master / rpl combination.  What
  remains a restatement of Chainer v3.5 inside the stand-in (a few lines each, upstream
  path cited there): Linear, F.lstm / L.LSTM, StatefulGRU, StatefulPeepholeLSTM,
  StatefulZoneoutLSTM, Convolution2D and the elementwise functions; the LSTM and no-reset
  MGRU algebra is additionally cross-checked against ``torch.nn.LSTMCell`` and an fp64
  evaluation in ``tests/test_oracle_cells.py``.
* bidirectional stacks: PARITY UNPINNED -- the reference has no such class;
  the definition is this module's (``birnn_forward_utterance``).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------
# Feature preparation (A1-A3)
# ----------------------------------------------------------------------------------------
def load_kaldi_feature_transform(filename):
    """scripts/util/kw_nn_utils.py:4-11 -- text nnet1 <Splice>/<AddShift>/<Rescale>."""
    with open(filename) as fid:
        s = fid.readlines()
    ft = {}
    ft["shape"] = [int(t) for t in s[1].split()[1:]]
    ft["shifts"] = [int(t) for t in s[2].split()[1:-1]]
    ft["addShift"] = np.asarray([float(t) for t in s[4].split()[3:-1]], dtype=F32)
    ft["rescale"] = np.asarray([float(t) for t in s[6].split()[3:-1]], dtype=F32)
    assert ft["addShift"].shape == (ft["shape"][0],)
    assert ft["rescale"].shape == (ft["shape"][0],)
    return ft


def select_transform_for_network(ft, network, splice=0, recurrent=None):
    """predict_folds.py:170-188 / evaluate.py:143-161: recurrent nets keep the shift-0
    block only, TDNN tiles that block ``winlen`` times, FF keeps the full vector."""
    if ft is None:
        return None
    ft = {k: (v.copy() if isinstance(v, np.ndarray) else list(v)) for k, v in ft.items()}
    if recurrent is None:
        recurrent = is_nn_recurrent(network)
    if recurrent:
        dim = ft["shape"][1]
        zi = ft["shifts"].index(0)
        ft["rescale"] = ft["rescale"][zi * dim:(zi + 1) * dim]
        ft["addShift"] = ft["addShift"][zi * dim:(zi + 1) * dim]
        ft["shape"][0] = dim
        ft["shifts"] = [0]
    elif network == "tdnn":
        dim = ft["shape"][1]
        zi = ft["shifts"].index(0)
        winlen = 2 * splice + 1
        ft["rescale"] = np.tile(ft["rescale"][zi * dim:(zi + 1) * dim], winlen)
        ft["addShift"] = np.tile(ft["addShift"][zi * dim:(zi + 1) * dim], winlen)
        ft["shape"][0] = dim * winlen
        ft["shifts"] = list(range(-splice, splice + 1))
    return ft


def apply_kaldi_feature_transform(x, ft):
    """kw_nn_utils.py:13-17 -- add then multiply, two separate fp32 roundings."""
    x1 = x + ft["addShift"]
    return x1 * ft["rescale"]


def splicing(data, shifts):
    """kw_utils.py:24-36 -- whole-array splice; the clamp is at the ends of the WHOLE
    array, so context bleeds across utterance boundaries."""
    data = np.asarray(data, dtype=F32)
    n, m = data.shape
    shifts = list(shifts)
    out = np.empty((n, m * len(shifts)), dtype=F32)
    idx = np.arange(n)
    for wi, w in enumerate(shifts):
        src = np.clip(idx + w, 0, n - 1)
        out[:, wi * m:(wi + 1) * m] = data[src]
    return out


def prepare_batch(data, idxs, winlen):
    """kw_nn_utils.py:19-43 (feature half) -- splice the frames ``idxs`` (sorted in place
    by the reference) with a window of ``winlen`` frames, clamped at the array ends."""
    data = np.asarray(data, dtype=F32)
    winhalf = int(winlen / 2)
    idxs = np.sort(np.asarray(idxs))
    num, dim = data.shape
    out = np.empty((len(idxs), winlen * dim), dtype=F32)
    for wi, w in enumerate(range(-winhalf, winhalf + 1)):
        src = np.clip(idxs + w, 0, num - 1)
        out[:, wi * dim:(wi + 1) * dim] = data[src]
    return out


def apply_time_delay_x(x, offsets, timedelay):
    """orcus_util.py:13-43 (x half, timedelay > 0): every utterance is right-padded with
    ``timedelay`` copies of its last frame."""
    assert timedelay > 0
    n_utt = len(offsets) - 1
    out = np.zeros((x.shape[0] + n_utt * timedelay, x.shape[1]), dtype=F32)
    new_off = np.array(offsets).copy()
    ptr = 0
    for o in range(n_utt):
        seg = x[offsets[o]:offsets[o + 1]]
        nxt = ptr + len(seg) + timedelay
        out[ptr:nxt] = np.pad(seg, ((0, timedelay), (0, 0)), "edge")
        new_off[o] = ptr
        ptr = nxt
    new_off[-1] = ptr
    return out, new_off


# ----------------------------------------------------------------------------------------
# Head (A11) and .lab writer (A13)
# ----------------------------------------------------------------------------------------
def logsum(lp, axis=1):
    """kw_utils.py:38-43 (used with axis=1 everywhere on the path)."""
    inf = 1e20
    lp = np.asarray(lp)
    mx = np.max(lp, axis=axis).reshape([lp.shape[0], 1])
    with np.errstate(all="ignore"):
        lps = mx + np.log(np.sum(np.exp(lp - mx), axis=axis)).reshape([lp.shape[0], 1])
    lps[np.isnan(lps)] = -inf
    return lps


def log_softmax(y):
    """predict_folds.py:57,88 -- ``y - logsum(y, axis=1)``."""
    return y - logsum(y, axis=1)


def head(y, ap=None):
    """evaluateModelForTest.py:75-77,110-112 -- optional ``y - ap`` then log-softmax."""
    if ap is not None:
        y = y - ap
    return y - logsum(y, axis=1)


def save_bin(filename, x):
    """kw_utils.py:4-12 -- uint32 rows, uint32 cols, raw row-major payload; the reader is
    recog_src/source/data.cpp:23-44 (int32 rows, int32 cols, float32 data)."""
    x = np.asarray(x)
    dims = np.asarray(x.shape, dtype=np.uint32)
    if len(dims) == 1:
        dims = np.resize(dims, 2)
        dims[1] = 1
    with open(filename, "wb") as fid:
        dims.tofile(fid, sep="")
        x.tofile(fid, sep="")


def load_bin(filename, dtype=F32):
    """kw_utils.py:14-22."""
    with open(filename, "rb") as fid:
        dims = np.fromfile(fid, dtype=np.uint32, sep="", count=2)
        x = np.fromfile(fid, dtype=dtype, sep="")
    return x.reshape(dims) if dims[1] > 1 else x


# ----------------------------------------------------------------------------------------
# Cells (A4-A9, Appendix A of SURVEY.md).  ``p`` is a dict of Chainer-layout parameters
# keyed WITHOUT the ``predictor/`` prefix, e.g. ``layer_0/W``, ``layer_0/upward/W``.
# ----------------------------------------------------------------------------------------
def is_nn_recurrent(n):
    """chainer_networks.py:186-187."""
    return n.endswith("lstm") or n.startswith("gru") or n.startswith("mgru")


def linear(x, w, b=None):
    """Chainer L.Linear / F.linear: ``y = x.dot(W.T) + b`` with W stored (out, in)."""
    y = x.dot(w.T)
    if b is not None:
        y = y + b
    return y.astype(x.dtype, copy=False)


def sigmoid(x):
    """Chainer's sigmoid forward: ``tanh(x * 0.5) * 0.5 + 0.5``."""
    half = x.dtype.type(0.5)
    return np.tanh(x * half) * half + half


_ACT = {
    "relu": lambda v: np.maximum(v, v.dtype.type(0)),
    "sigmoid": sigmoid,
    "tanh": np.tanh,
    "identity": lambda v: v,
}


def activation(name):
    return _ACT[name]


def mlp_forward(p, x, layers, act="relu"):
    """chainer_networks.py:19-22; dropout is the identity at inference."""
    f = _ACT[act]
    for l in range(layers):
        x = f(linear(x, p[f"layer_{l}/W"], p[f"layer_{l}/b"]))
    return linear(x, p["out/W"], p["out/b"])


def tdnn_forward(p, x, ksize, act="relu"):
    """chainer_networks.py:38-42 incl. the un-transposed reshape (quirk Q3): the flat
    (B, win*D) row is read as (B, C=D', 1, win) with c = n // win, w = n % win."""
    f = _ACT[act]
    win = sum(ksize) - len(ksize) + 1
    h = x.reshape(x.shape[0], -1, 1, win)
    for l, k in enumerate(ksize):
        w = p[f"layer_{l}/W"]  # (out, in, 1, k)
        b = p[f"layer_{l}/b"]
        wout = h.shape[3] - k + 1
        # valid cross-correlation along the last axis
        cols = np.stack([h[:, :, 0, j:j + wout] for j in range(k)], axis=-1)  # (B, in, wout, k)
        y = np.einsum("biwk,oik->bow", cols, w[:, :, 0, :]).astype(x.dtype)
        y = y + b[None, :, None]
        h = f(y)[:, :, None, :]
    h = h.reshape(h.shape[0], -1)
    return linear(h, p["out/W"], p["out/b"])


def lstm_step(p, prefix, x, h, c):
    """Chainer L.LSTM + F.lstm: gates = upward(x) [+ lateral(h)], 4H axis interleaved
    unit-major / gate-minor in the order a, i, f, o; lateral has no bias."""
    g = linear(x, p[prefix + "upward/W"], p[prefix + "upward/b"])
    if h is not None:
        g = g + linear(h, p[prefix + "lateral/W"])
    bsz = g.shape[0]
    hdim = g.shape[1] // 4
    g4 = g.reshape(bsz, hdim, 4)
    a = np.tanh(g4[:, :, 0])
    i = sigmoid(g4[:, :, 1])
    f = sigmoid(g4[:, :, 2])
    o = sigmoid(g4[:, :, 3])
    if c is None:
        c = np.zeros((bsz, hdim), dtype=x.dtype)
    c_new = a * i + f * c
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def peephole_lstm_step(p, prefix, x, h, c):
    """Chainer L.StatefulPeepholeLSTM: full-matrix peepholes on c_prev (i, f) and on
    c_new (o); gate layout as lstm_step."""
    g = linear(x, p[prefix + "upward/W"], p[prefix + "upward/b"])
    if h is not None:
        g = g + linear(h, p[prefix + "lateral/W"])
    bsz = g.shape[0]
    hdim = g.shape[1] // 4
    if c is None:
        c = np.zeros((bsz, hdim), dtype=x.dtype)
    g4 = g.reshape(bsz, hdim, 4)
    a = np.tanh(g4[:, :, 0])
    i = sigmoid(g4[:, :, 1] + linear(c, p[prefix + "peep_i/W"]))
    f = sigmoid(g4[:, :, 2] + linear(c, p[prefix + "peep_f/W"]))
    c_new = a * i + f * c
    o = sigmoid(g4[:, :, 3] + linear(c_new, p[prefix + "peep_o/W"]))
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def mgru_step(p, prefix, x, h, use_reset_gate, act):
    """scripts/common/MGRU.py:67-85 (StatefulMGRU.__call__).  Chainer's L.GRU
    (StatefulGRU) is the same step with use_reset_gate=True, act=tanh.  On the first step
    (h is None) every U_* term INCLUDING its bias is skipped and r is not computed."""
    f = _ACT[act]
    z = linear(x, p[prefix + "W_z/W"], p[prefix + "W_z/b"])
    h_bar = linear(x, p[prefix + "W/W"], p[prefix + "W/b"])
    if h is not None:
        z = z + linear(h, p[prefix + "U_z/W"], p[prefix + "U_z/b"])
        if use_reset_gate:
            r = sigmoid(linear(x, p[prefix + "W_r/W"], p[prefix + "W_r/b"])
                        + linear(h, p[prefix + "U_r/W"], p[prefix + "U_r/b"]))
            h_bar = h_bar + linear(r * h, p[prefix + "U/W"], p[prefix + "U/b"])
        else:
            h_bar = h_bar + linear(h, p[prefix + "U/W"], p[prefix + "U/b"])
    z = sigmoid(z)
    h_bar = f(h_bar)
    if h is not None:
        one = x.dtype.type(1)
        h_new = z * h_bar + (one - z) * h  # F.linear_interpolate(p, x, y) = p*x + (1-p)*y
    else:
        h_new = z * h_bar
    return h_new


_GRU_KIND = {  # chainer_networks.py:176-181
    "gru": (True, "tanh"),
    "mgrurelu": (False, "relu"),
    "mgrurelur": (True, "relu"),
}


class RecurrentNet:
    """Stateful per-time-step forward of the recurrent model specs
    (chainer_networks.py:44-161): x -> layers -> out Linear, raw logits."""

    def __init__(self, p, network, layers, bidirectional=False):
        self.p = p
        self.network = network
        self.layers = layers
        self.state = None
        self.reset_state()

    def reset_state(self):
        self.state = [(None, None) for _ in range(self.layers)]

    def __call__(self, x):
        p = self.p
        for l in range(self.layers):
            pre = f"layer_{l}/"
            h, c = self.state[l]
            if self.network in ("lstm", "zoneoutlstm", "zoneoutdropoutlstm"):
                h, c = lstm_step(p, pre, x, h, c)
            elif self.network == "peepholelstm":
                h, c = peephole_lstm_step(p, pre, x, h, c)
            else:
                reset, act = _GRU_KIND[self.network]
                h = mgru_step(p, pre, x, h, reset, act)
            self.state[l] = (h, c)
            x = h
        return linear(x, p["out/W"], p["out/b"])


def rnn_forward_utterance(p, network, layers, x):
    """Run one utterance (T, D) through a freshly reset recurrent net -> (T, C) logits."""
    net = RecurrentNet(p, network, layers)
    return np.concatenate([net(x[t:t + 1]) for t in range(x.shape[0])], axis=0)


def birnn_forward_utterance(p, network, layers, x):
    """Bidirectional stack (SURVEY A9 -- NOT in the reference; semantics defined by this
    build): per layer a forward cell ``layer_{l}/fwd/`` over t=0..T-1 and a backward cell
    ``layer_{l}/bwd/`` over t=T-1..0 on the same input, outputs concatenated [fwd, bwd];
    ``out`` is Linear(2H -> C).  network in {"lstm", "gru"}."""
    seq = x
    for l in range(layers):
        outs = []
        for d, order in (("fwd", range(seq.shape[0])), ("bwd", range(seq.shape[0] - 1, -1, -1))):
            pre = f"layer_{l}/{d}/"
            h = c = None
            hdim = p[pre + ("lateral/W" if network == "lstm" else "U/W")].shape[1]
            o = np.zeros((seq.shape[0], hdim), dtype=x.dtype)
            for t in order:
                if network == "lstm":
                    h, c = lstm_step(p, pre, seq[t:t + 1], h, c)
                else:
                    h = mgru_step(p, pre, seq[t:t + 1], h, *_GRU_KIND[network])
                o[t] = h[0]
            outs.append(o)
        seq = np.concatenate(outs, axis=1)
    return linear(seq, p["out/W"], p["out/b"])


def rpl4(p, h):
    """scripts/common/RPL.py:68-74: log-softmax, affine per class, logaddexp with lb."""
    x = log_softmax(h)
    g = x + x * p["W"] + p["b"]
    mx = np.maximum(g, p["lb"])
    mn = np.minimum(g, p["lb"])
    return mx + np.log(h.dtype.type(1.0) + np.exp(mn - mx))


def nn_with_rpl(master, folds, rpl, x):
    """scripts/common/evaluate.py:35-51 -- ensemble of RAW LOGITS.  ``master`` and
    ``folds`` are callables x -> logits; ``rpl`` is a callable or None."""
    k = len(folds)
    if master is not None and k == 0:
        h = master(x)
    elif master is not None:
        h = master(x) * F32(k)
        for f in folds:
            h = h + f(x)
        h = h / F32(2 * k)
    else:
        h = 0
        for f in folds:
            h = h + f(x)
        h = h / F32(k)
    if rpl is not None:
        h = rpl(h)
    return h


# ----------------------------------------------------------------------------------------
# predict() (A10) and the forward half of evaluateModelTestTri (A13)
# ----------------------------------------------------------------------------------------
def predict(model, x, offsets, network, winlen, timedelay, ft, recurrent=None):
    """scripts/common/predict_folds.py:27-95.  ``model`` is a callable (B, D) -> logits;
    recurrent models additionally expose ``reset_state()`` and are stateful per step."""
    if recurrent is None:
        recurrent = is_nn_recurrent(network)
    if recurrent:
        offsets = np.asarray(offsets)
        utt_len = offsets[1:] - offsets[:-1]
        utt_idx = np.flip(utt_len.argsort(), axis=0)
        utt_idx_rev = np.zeros(len(utt_len), dtype=np.int64)
        utt_idx_rev[utt_idx] = range(len(utt_len))
        xb = np.zeros((len(utt_len), utt_len[utt_idx[0]] + timedelay, x.shape[1]), dtype=F32)
        for i, idx in enumerate(utt_idx):
            x_ = x[offsets[idx]:offsets[idx + 1], :]
            x_ = np.pad(x_, ((0, timedelay), (0, 0)), mode="edge")
            if ft is not None:
                xb[i, :x_.shape[0], :] = apply_kaldi_feature_transform(x_, ft)
            else:
                xb[i, :x_.shape[0], :] = x_
        yb = None
        model.reset_state()
        for t in range(xb.shape[1]):
            batch_size = int(np.sum(utt_len > t))
            y = model(xb[:, t, :])
            y = y - logsum(y, axis=1)
            if yb is None:
                yb = np.zeros((xb.shape[0], xb.shape[1] - timedelay, y.shape[1]), dtype=F32)
            if t >= timedelay:
                yb[:batch_size, t - timedelay, :] = y[:batch_size]
        y_out = []
        for i, idx in enumerate(utt_idx_rev):
            y_out.append(yb[idx, :utt_len[i], :].reshape((utt_len[i], -1)))
        return np.concatenate(y_out, axis=0)
    y_out = []
    batch_size = 1024
    offset = 0
    while offset < x.shape[0]:
        offset_end = min(offset + batch_size, x.shape[0])
        x_ = prepare_batch(x, np.arange(offset, offset_end), winlen)
        if ft is not None:
            x_ = apply_kaldi_feature_transform(x_, ft)
        y = model(x_)
        y_out.append(y - logsum(y, axis=1))
        offset += batch_size
    return np.concatenate(y_out, axis=0)


def evaluate_forward(model, data, offsets, ap=None, rnn=False):
    """scripts/util/evaluateModelForTest.py:52-122 up to (not including) the decoder:
    returns the list of per-utterance (L_i, C) arrays that the reference hands to saveBin.
    The RNN branch applies NO time-delay compensation (quirk Q4)."""
    offsets = np.asarray(offsets)
    n_utt = len(offsets) - 1
    if rnn:
        lens = np.diff(offsets)
        order = np.flip(lens.argsort(), axis=0)
        rev = np.zeros(n_utt, dtype=np.int64)
        rev[order] = range(n_utt)
        xb = np.zeros((n_utt, lens[order[0]], data.shape[1]), dtype=F32)
        for i, idx in enumerate(order):
            xb[i, :lens[idx], :] = data[offsets[idx]:offsets[idx + 1], :]
        yb = None
        model.reset_state()
        for t in range(xb.shape[1]):
            bs = int(np.sum(lens > t))
            y = head(model(xb[:, t, :]), ap)
            if yb is None:
                yb = np.zeros((xb.shape[0], xb.shape[1], y.shape[1]), dtype=F32)
            yb[:bs, t, :] = y[:bs]
        return [yb[rev[i], :lens[i], :] for i in range(n_utt)]
    return [head(model(data[offsets[i]:offsets[i + 1], :]), ap) for i in range(n_utt)]


# ----------------------------------------------------------------------------------------
# Deterministic synthetic weights / workloads
# ----------------------------------------------------------------------------------------
def _lecun(rng, out_dim, in_dim):
    return (rng.standard_normal((out_dim, in_dim)) / np.sqrt(in_dim)).astype(F32)


def init_mlp(rng, in_dim, units, layers, n_out):
    p = {}
    d = in_dim
    for l in range(layers):
        p[f"layer_{l}/W"] = _lecun(rng, units, d)
        p[f"layer_{l}/b"] = np.zeros(units, dtype=F32)
        d = units
    p["out/W"] = _lecun(rng, n_out, d)
    p["out/b"] = np.zeros(n_out, dtype=F32)
    return p


def _init_lstm_cell(rng, p, pre, in_dim, units, forget_bias=1.0, peephole=False):
    p[pre + "upward/W"] = _lecun(rng, 4 * units, in_dim)
    b = np.zeros(4 * units, dtype=F32)
    b[2::4] = forget_bias  # rows 4j+2 = forget gate
    p[pre + "upward/b"] = b
    p[pre + "lateral/W"] = _lecun(rng, 4 * units, units)
    if peephole:
        for g in ("peep_i", "peep_f", "peep_o"):
            p[pre + g + "/W"] = _lecun(rng, units, units)


def _init_gru_cell(rng, p, pre, in_dim, units, use_reset_gate, bias_scale=0.0):
    names = ["W_z", "U_z", "W", "U"] + (["W_r", "U_r"] if use_reset_gate else [])
    for n in names:
        d = in_dim if n.startswith("W") else units
        p[pre + n + "/W"] = _lecun(rng, units, d)
        p[pre + n + "/b"] = (bias_scale * rng.standard_normal(units)).astype(F32)


def init_recurrent(rng, network, in_dim, units, layers, n_out, bidirectional=False, bias_scale=0.0):
    """Chainer default initialisers: LeCunNormal weights, zero biases, LSTM forget-gate
    bias 1 (plain ``L.LSTM`` only; zoneout/peephole links use plain Linear init).
    ``bias_scale`` > 0 randomises GRU biases so that the first-step "no U bias" rule is
    observable in tests."""
    p = {}
    d = in_dim
    for l in range(layers):
        dirs = ["fwd/", "bwd/"] if bidirectional else [""]
        for dd in dirs:
            pre = f"layer_{l}/{dd}"
            if network in ("lstm", "zoneoutlstm", "zoneoutdropoutlstm", "peepholelstm"):
                _init_lstm_cell(rng, p, pre, d, units,
                                forget_bias=1.0 if network == "lstm" else 0.0,
                                peephole=network == "peepholelstm")
            else:
                _init_gru_cell(rng, p, pre, d, units, _GRU_KIND[network][0], bias_scale)
        d = units * (2 if bidirectional else 1)
    p["out/W"] = _lecun(rng, n_out, d)
    p["out/b"] = np.zeros(n_out, dtype=F32)
    return p


def synth_lengths(rng, n_utt, total=None):
    """TIMIT-shaped utterance lengths: lognormal(ln 295, 0.28) clipped to [90, 780];
    optionally trimmed/extended so that sum == total."""
    ln = np.clip(np.round(rng.lognormal(np.log(295.0), 0.28, n_utt)), 90, 780).astype(np.int64)
    if total is not None:
        diff = int(total - ln.sum())
        i = 0
        while diff != 0:
            step = int(np.sign(diff)) * min(abs(diff), 16)
            new = int(np.clip(ln[i % n_utt] + step, 90, 780))
            diff -= new - ln[i % n_utt]
            ln[i % n_utt] = new
            i += 1
    return ln


def synth_set(seed, n_utt, dim=40, ivec_dim=0, total=None, utts_per_speaker=8):
    """Synthetic data set in the generate_folds.py:98-112 file convention:
    x (N, dim) f32, offsets (U+1,) int32 with a leading 0, ivectors (N, I) f32 or None."""
    rng = np.random.default_rng(seed)
    ln = synth_lengths(rng, n_utt, total)
    offsets = np.concatenate([[0], np.cumsum(ln)]).astype(np.int32)
    n = int(offsets[-1])
    x = rng.standard_normal((n, dim), dtype=F32)
    iv = None
    if ivec_dim:
        n_spk = (n_utt + utts_per_speaker - 1) // utts_per_speaker
        spk = (0.5 * rng.standard_normal((n_spk, ivec_dim))).astype(F32)
        iv = np.repeat(spk[np.arange(n_utt) // utts_per_speaker], ln, axis=0)
    return x, offsets, iv
