"""Independent sanity checks of the Chainer-cell restatements (parity unpinned by the
reference; see oracle/nnam_oracle.py header)."""
import numpy as np
import torch

from oracle import nnam_oracle as O


def _deinterleave_to_torch(w4h):
    """Chainer row 4j+k (k: a,i,f,o) -> torch LSTMCell blocks [i, f, g(=a), o]."""
    h = w4h.shape[0] // 4
    a, i, f, o = (w4h[k::4] for k in range(4))
    return np.concatenate([i, f, a, o], axis=0)


def test_lstm_step_matches_torch_lstmcell():
    rng = np.random.default_rng(0)
    p = O.init_recurrent(rng, "lstm", 12, 8, 1, 5)
    p["layer_0/upward/b"] = rng.standard_normal(32).astype(np.float32)
    cell = torch.nn.LSTMCell(12, 8)
    with torch.no_grad():
        cell.weight_ih.copy_(torch.from_numpy(_deinterleave_to_torch(p["layer_0/upward/W"])))
        cell.weight_hh.copy_(torch.from_numpy(_deinterleave_to_torch(p["layer_0/lateral/W"])))
        cell.bias_ih.copy_(torch.from_numpy(_deinterleave_to_torch(p["layer_0/upward/b"])))
        cell.bias_hh.zero_()
    x = rng.standard_normal((7, 3, 12)).astype(np.float32)
    h = c = None
    th = tc = torch.zeros(3, 8)
    for t in range(7):
        h, c = O.lstm_step(p, "layer_0/", x[t], h, c)
        with torch.no_grad():
            th, tc = cell(torch.from_numpy(x[t]), (th, tc))
        assert np.abs(h - th.numpy()).max() < 2e-6
        assert np.abs(c - tc.numpy()).max() < 2e-6


def test_mgru_no_reset_matches_closed_form_fp64():
    rng = np.random.default_rng(1)
    p = O.init_recurrent(rng, "mgrurelu", 6, 4, 1, 3, bias_scale=0.3)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    x = rng.standard_normal((5, 2, 6))
    h = None
    hh = None
    for t in range(5):
        h = O.mgru_step(p, "layer_0/", x[t].astype(np.float32), h, False, "relu")
        xz = x[t] @ p64["layer_0/W_z/W"].T + p64["layer_0/W_z/b"]
        xh = x[t] @ p64["layer_0/W/W"].T + p64["layer_0/W/b"]
        if hh is not None:
            xz = xz + hh @ p64["layer_0/U_z/W"].T + p64["layer_0/U_z/b"]
            xh = xh + hh @ p64["layer_0/U/W"].T + p64["layer_0/U/b"]
        z = 1 / (1 + np.exp(-xz))
        cand = np.maximum(xh, 0)
        hh = z * cand if hh is None else z * cand + (1 - z) * hh
        assert np.abs(h - hh).max() < 1e-5


def test_gru_first_step_skips_u_biases_and_reset_precedes_matmul():
    rng = np.random.default_rng(2)
    p = O.init_recurrent(rng, "gru", 6, 4, 1, 3, bias_scale=0.5)
    x = rng.standard_normal((2, 6)).astype(np.float32)
    h1 = O.mgru_step(p, "layer_0/", x, None, True, "tanh")
    z = O.sigmoid(x @ p["layer_0/W_z/W"].T + p["layer_0/W_z/b"])
    hb = np.tanh(x @ p["layer_0/W/W"].T + p["layer_0/W/b"])
    assert np.allclose(h1, z * hb, atol=1e-6)
    # h = 0 is NOT the same as h = None when the U biases are non-zero
    h0 = O.mgru_step(p, "layer_0/", x, np.zeros((2, 4), np.float32), True, "tanh")
    assert np.abs(h0 - h1).max() > 1e-3
    # second step: r is applied BEFORE the U matmul (MGRU.py:73-74)
    h2 = O.mgru_step(p, "layer_0/", x, h1, True, "tanh")
    r = O.sigmoid(x @ p["layer_0/W_r/W"].T + p["layer_0/W_r/b"] + h1 @ p["layer_0/U_r/W"].T + p["layer_0/U_r/b"])
    hb2 = np.tanh(x @ p["layer_0/W/W"].T + p["layer_0/W/b"] + (r * h1) @ p["layer_0/U/W"].T + p["layer_0/U/b"])
    z2 = O.sigmoid(x @ p["layer_0/W_z/W"].T + p["layer_0/W_z/b"] + h1 @ p["layer_0/U_z/W"].T + p["layer_0/U_z/b"])
    assert np.allclose(h2, z2 * hb2 + (1 - z2) * h1, atol=1e-6)


def test_zoneout_equals_lstm_at_inference_and_peephole_reduces_to_lstm():
    rng = np.random.default_rng(3)
    p = O.init_recurrent(rng, "peepholelstm", 5, 4, 1, 3)
    x = rng.standard_normal((6, 2, 5)).astype(np.float32)
    for k in ("peep_i", "peep_f", "peep_o"):
        p[f"layer_0/{k}/W"][:] = 0
    h = c = hp = cp = None
    for t in range(6):
        h, c = O.lstm_step(p, "layer_0/", x[t], h, c)
        hp, cp = O.peephole_lstm_step(p, "layer_0/", x[t], hp, cp)
        assert np.array_equal(h, hp) and np.array_equal(c, cp)


def test_tdnn_matches_explicit_loop():
    rng = np.random.default_rng(4)
    ksize = [3, 2]
    win = sum(ksize) - len(ksize) + 1  # 4
    d = 5
    p = {
        "layer_0/W": rng.standard_normal((6, d, 1, 3)).astype(np.float32),
        "layer_0/b": rng.standard_normal(6).astype(np.float32),
        "layer_1/W": rng.standard_normal((7, 6, 1, 2)).astype(np.float32),
        "layer_1/b": rng.standard_normal(7).astype(np.float32),
        "out/W": rng.standard_normal((3, 7)).astype(np.float32),
        "out/b": rng.standard_normal(3).astype(np.float32),
    }
    x = rng.standard_normal((2, win * d)).astype(np.float32)
    got = O.tdnn_forward(p, x, ksize, "relu")
    h = x.reshape(2, d, win).astype(np.float64)  # quirk Q3: (B, C, W) straight from the flat row
    for l, k in enumerate(ksize):
        w = p[f"layer_{l}/W"][:, :, 0, :].astype(np.float64)
        wo = h.shape[2] - k + 1
        y = np.zeros((2, w.shape[0], wo))
        for o in range(w.shape[0]):
            for t in range(wo):
                y[:, o, t] = (h[:, :, t:t + k] * w[o]).sum(axis=(1, 2)) + p[f"layer_{l}/b"][o]
        h = np.maximum(y, 0)
    want = h.reshape(2, -1) @ p["out/W"].T.astype(np.float64) + p["out/b"]
    assert np.abs(got - want).max() < 1e-4


def test_predict_rnn_quirk_q4_and_ff_chunking():
    rng = np.random.default_rng(5)
    p = O.init_recurrent(rng, "lstm", 4, 6, 2, 7)
    net = O.RecurrentNet(p, "lstm", 2)
    offsets = np.array([0, 9, 15, 28], dtype=np.int32)
    x = rng.standard_normal((28, 4)).astype(np.float32)
    td = 3
    y = O.predict(net, x, offsets, "lstm", 1, td, None)
    assert y.shape == (28, 7)
    for u in range(3):
        seg = y[offsets[u]:offsets[u + 1]]
        assert np.all(seg[-td:] == 0.0)  # Q4: last `timedelay` frames stay zero
        xp = np.pad(x[offsets[u]:offsets[u + 1]], ((0, td), (0, 0)), mode="edge")
        full = O.log_softmax(O.rnn_forward_utterance(p, "lstm", 2, xp))
        assert np.allclose(seg[:-td], full[td:len(seg)], atol=1e-5)
    pm = O.init_mlp(rng, 4 * 3, 8, 2, 7)
    big = rng.standard_normal((2500, 4)).astype(np.float32)
    yf = O.predict(lambda v: O.mlp_forward(pm, v, 2), big, None, "ff", 3, 0, None)
    want = O.log_softmax(O.mlp_forward(pm, O.splicing(big, range(-1, 2)), 2))
    assert np.allclose(yf, want, atol=1e-5)


def test_ensemble_and_rpl4():
    rng = np.random.default_rng(6)
    y = [rng.standard_normal((4, 9)).astype(np.float32) for _ in range(3)]
    m = lambda v: y[0]
    folds = [lambda v: y[1], lambda v: y[2]]
    assert np.allclose(O.nn_with_rpl(m, [], None, None), y[0])
    assert np.allclose(O.nn_with_rpl(m, folds, None, None), (2 * y[0] + y[1] + y[2]) / 4, atol=1e-6)
    assert np.allclose(O.nn_with_rpl(None, folds, None, None), (y[1] + y[2]) / 2, atol=1e-6)
    prm = {"W": np.zeros((1, 9), np.float32), "b": np.zeros((1, 9), np.float32), "lb": np.full((1, 9), -20, np.float32)}
    out = O.rpl4(prm, y[0])
    assert np.allclose(out, np.logaddexp(O.log_softmax(y[0]), -20.0), atol=1e-6)


def test_synth_set_shapes():
    x, off, iv = O.synth_set(1, 40, ivec_dim=10, total=12000)
    assert off[0] == 0 and off[-1] == 12000 == x.shape[0] and off.dtype == np.int32
    assert iv.shape == (12000, 10)
    ln = np.diff(off)
    assert ln.min() >= 90 and ln.max() <= 780
    assert np.array_equal(iv[off[0]], iv[off[8] - 1]) and not np.array_equal(iv[off[0]], iv[off[8]])
