"""BASELINE.json's full-size configurations on the GPU: size-independent properties over the whole output plus oracle
parity on a random sample of frames / utterances (the oracle finishes a sample in seconds).  -m gpu."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nnam_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def nn():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import nnacousticmodeling_b200 as _nn
    return _nn


def _rows_are_log_distributions(y, tol=2e-4):
    lse = O.logsum(y.astype(np.float32), axis=1)
    return float(np.abs(lse).max()) < tol


def test_cfg2_full_train_shaped_set(nn, golden_dir):
    """configs[1]: 6x2048 MLP on 440 spliced fMLLR + 100-dim i-vectors, 3696 utterances / 1,124,823 frames."""
    x, off, iv = O.synth_set(1234, 3696, 40, 100, total=1124823)
    assert len(x) == 1124823 and off[-1] == 1124823
    ft = nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform"))
    p = O.init_mlp(np.random.default_rng(4321), 540, 2048, 6, 1909)
    m = nn.get_nn("ff", 6, [2048], 1909, nn.F.relu, [5])
    m.load_params(p)
    out = {}
    for mode in ("bf16", "fp32"):
        m.precision = mode
        out[mode] = nn.predict(m, x, None, 1909, "ff", 0, 11, 0, ft, progress=False, ivectors=iv)
        assert out[mode].shape == (1124823, 1909) and np.isfinite(out[mode]).all()
        # every row is a log-distribution: log(sum(exp(row))) == 0
        step = 7  # a strided 1/7 of the rows keeps the host check short
        assert _rows_are_log_distributions(out[mode][::step])
    # oracle parity on a random sample of frames, including both ends of the set (splice clamp, quirk Q1)
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([np.arange(8), np.arange(len(x) - 8, len(x)), rng.integers(0, len(x), 3000)]))
    oft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    feats = np.concatenate((O.apply_kaldi_feature_transform(O.prepare_batch(x, idx, 11), oft), iv[idx]), axis=1)
    want = O.log_softmax(O.mlp_forward(p, feats, 6))
    assert np.abs(out["fp32"][idx] - want).max() < 1e-3
    assert np.abs(out["bf16"][idx] - want).max() < 5e-2
    near = want[np.arange(len(idx)), out["bf16"][idx].argmax(axis=1)] >= want.max(axis=1) - 1e-2
    assert near.mean() >= 0.995


def test_cfg3_full_test_shaped_set(nn, golden_dir):
    """configs[2]: 4x512 LSTM on 40-dim fMLLR, timedelay 5, 1344 utterances (test-shaped)."""
    x, off, _ = O.synth_set(1237, 1344)
    p = O.init_recurrent(np.random.default_rng(4321), "lstm", 40, 512, 4, 1909)
    m = nn.get_nn("lstm", 4, [512], 1909, nn.F.relu, [5])
    m.load_params(p)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            "lstm", 0, True)
    got = {}
    for mode in ("bf16", "fp32"):
        m.precision = mode
        got[mode] = nn.predict(m, x, off, 1909, "lstm", 0, 1, 5, ft, progress=False)
        assert got[mode].shape == (off[-1], 1909) and np.isfinite(got[mode]).all()
    # quirk Q4: the last 5 frames of every utterance stay 0, every other row is a log-distribution
    tail = np.zeros(off[-1], bool)
    for u in range(1344):
        tail[off[u + 1] - 5:off[u + 1]] = True
    assert np.all(got["fp32"][tail] == 0) and np.all(got["bf16"][tail] == 0)
    assert _rows_are_log_distributions(got["fp32"][~tail][::5]) and _rows_are_log_distributions(got["bf16"][~tail][::5])
    # oracle parity on a sample of utterances: the shortest, the longest and a few random ones
    lens = np.diff(off)
    pick = sorted({int(lens.argmin()), int(lens.argmax()), *np.random.default_rng(1).integers(0, 1344, 4).tolist()})
    oft = O.select_transform_for_network(O.load_kaldi_feature_transform(
        os.path.join(golden_dir, "final.feature_transform")), "lstm")
    for u in pick:
        xs = x[off[u]:off[u + 1]]
        want = O.predict(O.RecurrentNet(p, "lstm", 4), xs, np.array([0, len(xs)]), "lstm", 1, 5, oft)
        assert np.abs(got["fp32"][off[u]:off[u + 1]] - want).max() < 1e-3
        assert np.abs(got["bf16"][off[u]:off[u + 1]] - want).max() < 5e-2
