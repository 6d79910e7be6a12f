"""BASELINE.json's configurations at FULL size on the GPU, every precision mode, against the fp32 oracle.  -m gpu.

The whole synthetic set goes through the product path (`predict()`); size-independent properties are checked over the
whole output and oracle parity on a large random sample of frames / utterances (the oracle finishes a sample in
seconds).  north_star's gates, asserted here exactly as stated -- RAW frame-argmax agreement, no near-tie allowance:

  fp32 mode  (bf16x3 GEMMs)        max |err| <= 1e-3
  16-bit throughput mode           max |err| <= 5e-2  AND  argmax agreement >= 99.5 %

Which 16-bit mode meets the second gate was MEASURED (tests/tools/gpu_parity_table.py, profiles/r02_parity_table.md): the
logits of random-init nets are almost flat (top-1/top-2 gaps of 1e-3 are common), single-pass bf16 (8-bit
significand) flips 1.1-1.4 % of the argmaxes -- 0.6-0.9 % even with bf16-exact weights on both sides -- and does NOT
meet it; single-pass fp16 (11-bit significand, same tensor-pipe rate) agrees on 99.8-99.9 % and does.  So "fp16" is
the throughput mode, asserted at >= 0.995; plain "bf16" keeps its max-abs assertion and a 0.98 regression floor, and
"bf16+a" (split activations) on bf16-exact weights is the bf16-strict mode that meets the gate at twice the passes.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nnam_oracle as O  # noqa: E402

GATE = 0.995


@pytest.fixture(scope="module")
def nn():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import nnacousticmodeling_b200 as _nn
    return _nn


def _rows_are_log_distributions(y, tol=2e-4):
    lse = O.logsum(y.astype(np.float32), axis=1)
    return float(np.abs(lse).max()) < tol


def _stats(tag, got, want):
    err = float(np.abs(got - want).max())
    agree = float(np.mean(got.argmax(axis=1) == want.argmax(axis=1)))
    print(f"[parity] {tag}: frames={len(want)} max|err|={err:.2e} raw argmax agreement={agree:.4f}")
    return err, agree


def _round16(p, dtype):
    return {k: (torch.from_numpy(v).to(dtype).float().numpy() if k.endswith("/W") else v) for k, v in p.items()}


def _check(tag, mode, got, want):
    err, agree = _stats(f"{tag} {mode}", got, want)
    if mode == "fp32":
        assert err < 1e-3 and agree >= GATE
    elif mode == "bf16":
        assert err < 5e-2 and agree >= 0.98  # bf16 does not meet the 99.5 % gate (module docstring); regression floor
    else:
        assert err < 5e-2 and agree >= GATE


def _ff_sample(x, iv, golden_dir, n=20000, seed=0):
    rng = np.random.default_rng(seed)
    idx = np.unique(np.concatenate([np.arange(8), np.arange(len(x) - 8, len(x)), rng.integers(0, len(x), n)]))
    oft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    feats = O.apply_kaldi_feature_transform(O.prepare_batch(x, idx, 11), oft)  # both array ends: splice clamp, quirk Q1
    if iv is not None:
        feats = np.concatenate((feats, iv[idx]), axis=1)
    return idx, feats


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_mlp_full_sets_every_mode(nn, golden_dir, cfg):
    """configs[0]: 6x1024 MLP on 440 spliced fMLLR, 1344 utterances; configs[1]: 6x2048 MLP on 440 + 100-dim i-vectors,
    3696 utterances / 1,124,823 frames."""
    if cfg == "cfg1":
        x, off, iv = O.synth_set(1234, 1344)
        units, d_in = 1024, 440
    else:
        x, off, iv = O.synth_set(1234, 3696, 40, 100, total=1124823)
        assert len(x) == 1124823 and off[-1] == 1124823
        units, d_in = 2048, 540
    ft = nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform"))
    p = O.init_mlp(np.random.default_rng(4321), d_in, units, 6, 1909)
    idx, feats = _ff_sample(x, iv, golden_dir)
    want = O.log_softmax(O.mlp_forward(p, feats, 6))
    m = nn.get_nn("ff", 6, [units], 1909, nn.F.relu, [5])
    m.load_params(p)
    out = nn.empty_pinned((len(x), 1909))
    for mode in ("fp16", "bf16", "fp32"):
        m.precision = mode
        nn.predict(m, x, None, 1909, "ff", 0, 11, 0, ft, progress=False, ivectors=iv, out=out)
        assert np.isfinite(out[::7]).all() and _rows_are_log_distributions(out[::7])  # a strided 1/7 of the rows
        _check(cfg, mode, out[idx], want)
    # the bf16-strict mode that meets the gate: identical (bf16-representable) weights on both sides, activations
    # carried as bf16 hi/lo pairs (two tensor passes)
    pe = _round16(p, torch.bfloat16)
    m.load_params(pe)
    m.precision = "bf16+a"
    nn.predict(m, x, None, 1909, "ff", 0, 11, 0, ft, progress=False, ivectors=iv, out=out)
    _check(cfg, "bf16+a (bf16-exact weights)", out[idx], O.log_softmax(O.mlp_forward(pe, feats, 6)))


def _utt_sample(off, n, seed=1):
    lens = np.diff(off)
    return sorted({int(lens.argmin()), int(lens.argmax()),
                   *np.random.default_rng(seed).integers(0, len(lens), n).tolist()})


def test_cfg3_lstm_full_test_shaped_set(nn, golden_dir):
    """configs[2]: 4x512 LSTM on 40-dim fMLLR, timedelay 5, 1344 utterances (test-shaped)."""
    x, off, _ = O.synth_set(1237, 1344)
    p = O.init_recurrent(np.random.default_rng(4321), "lstm", 40, 512, 4, 1909)
    m = nn.get_nn("lstm", 4, [512], 1909, nn.F.relu, [5])
    m.load_params(p)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            "lstm", 0, True)
    oft = O.select_transform_for_network(O.load_kaldi_feature_transform(
        os.path.join(golden_dir, "final.feature_transform")), "lstm")
    lens = np.diff(off)
    pick = _utt_sample(off, 40)
    # the reference's time-major loop over the sampled utterances (timedelay, quirk Q4)
    sub = np.concatenate([[0], np.cumsum(lens[pick])])
    y = O.predict(O.RecurrentNet(p, "lstm", 4), np.concatenate([x[off[u]:off[u + 1]] for u in pick]), sub, "lstm", 1, 5,
                  oft)
    rows = np.concatenate([np.arange(off[u], off[u + 1] - 5) for u in pick])
    want = np.concatenate([y[sub[i]:sub[i + 1] - 5] for i in range(len(pick))])
    tail = np.zeros(off[-1], bool)
    for u in range(1344):
        tail[off[u + 1] - 5:off[u + 1]] = True
    for mode in ("fp16", "bf16", "fp32"):
        m.precision = mode
        got = nn.predict(m, x, off, 1909, "lstm", 0, 1, 5, ft, progress=False)
        assert got.shape == (off[-1], 1909) and np.isfinite(got).all()
        # quirk Q4: the last 5 frames of every utterance stay 0, every other row is a log-distribution
        assert np.all(got[tail] == 0) and _rows_are_log_distributions(got[~tail][::5])
        _check("cfg3", mode, got[rows], want)


@pytest.mark.parametrize("net", ["blstm", "bgru"])
def test_cfg4_bidirectional_full_geometry(nn, golden_dir, net):
    """configs[3]: 4x(2x512) bidirectional LSTM / GRU on 40 fMLLR + 100 i-vectors -> 1909, 1344 utterances."""
    cell = "gru" if net == "bgru" else "lstm"
    x, off, iv = O.synth_set(1238, 1344, 40, 100)
    p = O.init_recurrent(np.random.default_rng(4321), cell, 140, 512, 4, 1909, bidirectional=True)
    m = nn.get_nn(net, 4, [512], 1909, nn.F.relu, [5])
    m.load_params(p)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            net, 0, True)
    oft = O.select_transform_for_network(O.load_kaldi_feature_transform(
        os.path.join(golden_dir, "final.feature_transform")), "lstm")
    pick = _utt_sample(off, 10)
    want = np.concatenate([O.log_softmax(O.birnn_forward_utterance(
        p, cell, 4, np.concatenate((O.apply_kaldi_feature_transform(x[off[u]:off[u + 1]], oft), iv[off[u]:off[u + 1]]),
                                   axis=1))) for u in pick])
    rows = np.concatenate([np.arange(off[u], off[u + 1]) for u in pick])
    for mode in ("fp16", "bf16", "fp32"):
        m.precision = mode
        got = nn.predict(m, x, off, 1909, net, 0, 1, 0, ft, progress=False, ivectors=iv)
        assert got.shape == (off[-1], 1909) and np.isfinite(got).all() and _rows_are_log_distributions(got[::5])
        _check(f"cfg4 {net}", mode, got[rows], want)


def test_cfg5_ten_fold_mlp_ensemble(nn, golden_dir):
    """configs[4], MLP half: logit mean of 10 fold 6x1024 MLPs (evaluate.py:35-51) on the test-shaped set."""
    x, off, _ = O.synth_set(1239, 1344)
    ft = nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform"))
    ps = [O.init_mlp(np.random.default_rng(5000 + k), 440, 1024, 6, 1909) for k in range(10)]
    ms = []
    for q in ps:
        m = nn.get_nn("ff", 6, [1024], 1909, nn.F.relu, [5])
        m.load_params(q)
        ms.append(m)
    idx, feats = _ff_sample(x, None, golden_dir, n=4000)
    want = O.log_softmax(O.nn_with_rpl(None, [(lambda v, q=q: O.mlp_forward(q, v, 6)) for q in ps], None, feats))
    head = nn.HeadSpec(weights=[0.1] * 10)
    for mode in ("fp16", "fp32"):
        for m in ms:
            m.precision = mode
        got = nn.predict(ms, x, None, 1909, "ff", 0, 11, 0, ft, progress=False, head=head)
        assert _rows_are_log_distributions(got[::7])
        _check("cfg5 10xMLP", mode, got[idx], want)


def test_cfg5b_ten_fold_blstm_ensemble(nn, golden_dir):
    """configs[4], BLSTM half: logit mean of 10 fold 4x(2x512) BLSTMs with i-vectors on the test-shaped set."""
    x, off, iv = O.synth_set(1240, 1344, 40, 100)
    ps = [O.init_recurrent(np.random.default_rng(6000 + k), "lstm", 140, 512, 4, 1909, bidirectional=True)
          for k in range(10)]
    ms = []
    for q in ps:
        m = nn.get_nn("blstm", 4, [512], 1909, nn.F.relu, [5])
        m.load_params(q)
        ms.append(m)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            "blstm", 0, True)
    oft = O.select_transform_for_network(O.load_kaldi_feature_transform(
        os.path.join(golden_dir, "final.feature_transform")), "lstm")
    pick = _utt_sample(off, 2)
    wants = []
    for u in pick:
        xs = np.concatenate((O.apply_kaldi_feature_transform(x[off[u]:off[u + 1]], oft), iv[off[u]:off[u + 1]]), axis=1)
        wants.append(O.log_softmax(sum(O.birnn_forward_utterance(q, "lstm", 4, xs) for q in ps) / np.float32(10)))
    want = np.concatenate(wants)
    rows = np.concatenate([np.arange(off[u], off[u + 1]) for u in pick])
    head = nn.HeadSpec(weights=[0.1] * 10)
    for mode in ("fp16", "fp32"):
        for m in ms:
            m.precision = mode
        got = nn.predict(ms, x, off, 1909, "blstm", 0, 1, 0, ft, progress=False, ivectors=iv, head=head)
        assert got.shape == (off[-1], 1909) and _rows_are_log_distributions(got[::5])
        _check("cfg5b 10xBLSTM", mode, got[rows], want)
