"""Recurrent path (K3 + schedule) against the oracle.  -m gpu."""
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


def _lstm(nn, seed, network, in_dim, units, layers, n_out, precision="fp32", bidirectional=False):
    p = O.init_recurrent(np.random.default_rng(seed), "lstm" if network == "blstm" else network, in_dim, units, layers,
                         n_out, bidirectional=bidirectional)
    rng = np.random.default_rng(seed + 100)
    for k in p:
        if k.endswith("upward/b") or k == "out/b":
            p[k] = p[k] + (0.1 * rng.standard_normal(p[k].shape)).astype(np.float32)
    m = nn.get_nn(network, layers, [units], n_out, nn.F.relu, [5])
    m.load_params(p)
    m.precision = precision
    return m, p


def _offsets(lens):
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


@pytest.mark.parametrize("units,layers,lens,td", [
    (64, 1, [5], 0),
    (64, 2, [7, 3, 12, 1, 9], 0),
    (128, 2, [30, 41, 17, 25, 33, 8] * 7, 3),   # 42 utterances: more than one batch of 32
    (512, 4, [90, 120, 75, 101], 5),            # BASELINE config 3 geometry (40 -> 4x512 -> 1909), timedelay 5
])
def test_predict_lstm_fp32_mode(nn, golden_dir, units, layers, lens, td):
    off = _offsets(lens)
    rng = np.random.default_rng(sum(lens))
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    n_out = 1909 if units == 512 else 39
    m, p = _lstm(nn, units + layers, "lstm", 40, units, layers, n_out)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            "lstm", 0, True)
    want = O.predict(O.RecurrentNet(p, "lstm", layers), x, off, "lstm", 1, td, ft)
    got = nn.predict(m, x, off, n_out, "lstm", 0, 1, td, ft, progress=False)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-3
    if td:
        for u in range(len(lens)):  # quirk Q4: the reference leaves the last `timedelay` frames at 0
            assert np.all(got[off[u + 1] - min(td, lens[u]):off[u + 1]] == 0)
        fixed = nn.predict(m, x, off, n_out, "lstm", 0, 1, td, ft, progress=False, fix_timedelay_tail=True)
        assert np.all(np.abs(fixed).sum(axis=1) > 0)


def test_predict_lstm_16bit_modes(nn):
    """fp16 = the 16-bit mode that meets north_star's gate (<= 5e-2, >= 99.5 % RAW argmax agreement); bf16 meets the
    max-abs tolerance only (0.98 is a regression floor; see tests/test_gpu_full_size.py)."""
    lens = [60, 85, 44, 70, 52, 66, 91, 38] * 4
    off = _offsets(lens)
    x = np.random.default_rng(1).standard_normal((off[-1], 40)).astype(np.float32)
    m, p = _lstm(nn, 3, "lstm", 40, 512, 4, 1909, precision="fp16")
    want = O.predict(O.RecurrentNet(p, "lstm", 4), x, off, "lstm", 1, 0, None)
    got = nn.predict(m, x, off, 1909, "lstm", 0, 1, 0, None, progress=False)
    assert np.abs(got - want).max() < 5e-2
    assert np.mean(got.argmax(axis=1) == want.argmax(axis=1)) >= 0.995
    m.precision = "bf16"
    got = nn.predict(m, x, off, 1909, "lstm", 0, 1, 0, None, progress=False)
    assert np.abs(got - want).max() < 5e-2
    assert np.mean(got.argmax(axis=1) == want.argmax(axis=1)) >= 0.98


def test_zoneout_lstm_is_lstm_arithmetic(nn):
    lens = [9, 4, 6]
    off = _offsets(lens)
    x = np.random.default_rng(2).standard_normal((off[-1], 40)).astype(np.float32)
    m, p = _lstm(nn, 4, "zoneoutlstm", 40, 64, 2, 39)
    want = O.predict(O.RecurrentNet(p, "zoneoutlstm", 2), x, off, "zoneoutlstm", 1, 0, None)
    got = nn.predict(m, x, off, 39, "zoneoutlstm", 0, 1, 0, None, progress=False)
    assert np.abs(got - want).max() < 1e-3


def test_blstm_with_ivectors(nn, golden_dir):
    """BASELINE config 4 geometry at small scale: bidirectional LSTM on 40 fMLLR + 100 i-vector."""
    lens = [33, 20, 41, 12, 27]
    off = _offsets(lens)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    iv = np.repeat((0.5 * rng.standard_normal((len(lens), 100))).astype(np.float32), lens, axis=0)
    m, p = _lstm(nn, 6, "blstm", 140, 128, 2, 1909, bidirectional=True)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            "blstm", 0, True)
    feats = np.concatenate((O.apply_kaldi_feature_transform(x, ft), iv), axis=1)
    want = np.concatenate([O.log_softmax(O.birnn_forward_utterance(p, "lstm", 2, feats[off[u]:off[u + 1]]))
                           for u in range(len(lens))])
    got = nn.predict(m, x, off, 1909, "blstm", 0, 1, 0, ft, progress=False, ivectors=iv)
    assert np.abs(got - want).max() < 1e-3


def test_stateful_call_and_reset(nn):
    m, p = _lstm(nn, 7, "lstm", 40, 64, 3, 39)
    ref = O.RecurrentNet(p, "lstm", 3)
    rng = np.random.default_rng(8)
    xs = rng.standard_normal((6, 5, 40)).astype(np.float32)
    m.reset_state()
    for t in range(6):
        assert np.abs(m(xs[t]) - ref(xs[t])).max() < 1e-3
    m.reset_state()
    ref.reset_state()
    assert np.abs(m(xs[0]) - ref(xs[0])).max() < 1e-3


def test_multi_shard_equals_single(nn):
    from nnacousticmodeling_b200 import recurrent_engine
    lens = [15, 22, 9, 31, 18, 12, 27, 20, 11]
    off = _offsets(lens)
    x = np.random.default_rng(9).standard_normal((off[-1], 40)).astype(np.float32)
    m, p = _lstm(nn, 10, "lstm", 40, 64, 2, 39)
    full = nn.predict(m, x, off, 39, "lstm", 0, 1, 2, None, progress=False)
    parts = np.zeros_like(full)
    for u0, u1 in nn.partition_utterances(off, 3):
        recurrent_engine.forward_utterances(m, x, off, parts, u0, u1, timedelay=2, device=0)
    assert np.array_equal(full, parts)


def _gru(nn, seed, network, in_dim, units, layers, n_out, precision="fp32", bidirectional=False):
    base = "gru" if network == "bgru" else network
    p = O.init_recurrent(np.random.default_rng(seed), base, in_dim, units, layers, n_out, bidirectional=bidirectional,
                         bias_scale=0.3)  # non-zero U biases: the "first step has no U terms" rule is observable
    m = nn.get_nn(network, layers, [units], n_out, nn.F.relu, [5])
    m.load_params(p)
    m.precision = precision
    return m, p


@pytest.mark.parametrize("network", ["gru", "mgrurelu", "mgrurelur"])
@pytest.mark.parametrize("units,layers,lens,td", [(64, 2, [7, 3, 12, 1, 9], 0), (128, 2, [30, 41, 17, 25, 33, 8] * 7, 3),
                                                  (512, 4, [60, 75, 44, 58], 5)])
def test_predict_gru_family_fp32_mode(nn, golden_dir, network, units, layers, lens, td):
    off = _offsets(lens)
    x = np.random.default_rng(sum(lens) + 1).standard_normal((off[-1], 40)).astype(np.float32)
    n_out = 1909 if units == 512 else 39
    m, p = _gru(nn, units + layers, network, 40, units, layers, n_out)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            network, 0, True)
    want = O.predict(O.RecurrentNet(p, network, layers), x, off, network, 1, td, ft)
    got = nn.predict(m, x, off, n_out, network, 0, 1, td, ft, progress=False)
    assert np.abs(got - want).max() < 1e-3


def test_gru_bf16_mode_and_bidirectional(nn):
    lens = [40, 55, 31, 47, 62, 28]
    off = _offsets(lens)
    x = np.random.default_rng(11).standard_normal((off[-1], 40)).astype(np.float32)
    m, p = _gru(nn, 12, "gru", 40, 512, 4, 1909, precision="bf16")
    want = O.predict(O.RecurrentNet(p, "gru", 4), x, off, "gru", 1, 0, None)
    got = nn.predict(m, x, off, 1909, "gru", 0, 1, 0, None, progress=False)
    assert np.abs(got - want).max() < 5e-2
    m.precision = "fp16"
    got = nn.predict(m, x, off, 1909, "gru", 0, 1, 0, None, progress=False)
    assert np.abs(got - want).max() < 1e-2 and np.mean(got.argmax(axis=1) == want.argmax(axis=1)) >= 0.995
    mb, pb = _gru(nn, 13, "bgru", 40, 128, 2, 39, bidirectional=True)
    want = np.concatenate([O.log_softmax(O.birnn_forward_utterance(pb, "gru", 2, x[off[u]:off[u + 1]]))
                           for u in range(len(lens))])
    got = nn.predict(mb, x, off, 39, "bgru", 0, 1, 0, None, progress=False)
    assert np.abs(got - want).max() < 1e-3


def test_gru_stateful_call(nn):
    m, p = _gru(nn, 14, "gru", 40, 64, 2, 39)
    ref = O.RecurrentNet(p, "gru", 2)
    xs = np.random.default_rng(15).standard_normal((5, 4, 40)).astype(np.float32)
    m.reset_state()
    for t in range(5):
        assert np.abs(m(xs[t]) - ref(xs[t])).max() < 1e-3


@pytest.mark.parametrize("units,layers,lens,td", [(64, 2, [7, 3, 12, 1, 9], 0), (128, 2, [30, 41, 17, 25, 33, 8] * 3, 2),
                                                  (512, 2, [40, 55, 31], 5)])
def test_predict_peephole_lstm(nn, golden_dir, units, layers, lens, td):
    """L.StatefulPeepholeLSTM (chainer_networks.py:103-121): full-matrix peepholes, time-step launches on the device."""
    off = _offsets(lens)
    x = np.random.default_rng(sum(lens) + 2).standard_normal((off[-1], 40)).astype(np.float32)
    n_out = 1909 if units == 512 else 39
    p = O.init_recurrent(np.random.default_rng(units), "peepholelstm", 40, units, layers, n_out)
    rng = np.random.default_rng(units + 1)
    for k in p:
        if k.endswith("upward/b"):
            p[k] = p[k] + (0.1 * rng.standard_normal(p[k].shape)).astype(np.float32)
    m = nn.get_nn("peepholelstm", layers, [units], n_out, nn.F.relu, [5])
    m.load_params(p)
    ft = nn.adapt_transform(nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform")),
                            "peepholelstm", 0, True)
    want = O.predict(O.RecurrentNet(p, "peepholelstm", layers), x, off, "peepholelstm", 1, td, ft)
    got = nn.predict(m, x, off, n_out, "peepholelstm", 0, 1, td, ft, progress=False)
    assert np.abs(got - want).max() < 1e-3
    for prec in ("bf16", "fp16"):
        m.precision = prec
        got16 = nn.predict(m, x, off, n_out, "peepholelstm", 0, 1, td, ft, progress=False)
        assert np.abs(got16 - want).max() < 5e-2
    # stateful per-step surface
    m.precision = "fp32"
    ref = O.RecurrentNet(p, "peepholelstm", layers)
    xs = np.random.default_rng(3).standard_normal((4, 6, 40)).astype(np.float32)
    m.reset_state()
    for t in range(4):
        assert np.abs(m(xs[t]) - ref(xs[t])).max() < 1e-3


@pytest.mark.parametrize("network", ["lstm", "gru", "peepholelstm"])
def test_timedelay_longer_than_some_utterances_and_empty_input(nn, network):
    """Edge cases of predict_folds.py:34-64: utterances shorter than the delay produce only zero rows (quirk Q4), and
    an empty feature matrix gives an empty output."""
    lens = [2, 9, 1, 4, 30]
    td = 3
    off = _offsets(lens)
    x = np.random.default_rng(5).standard_normal((off[-1], 40)).astype(np.float32)
    p = O.init_recurrent(np.random.default_rng(6), network, 40, 64, 2, 39, bias_scale=0.2)
    m = nn.get_nn(network, 2, [64], 39, nn.F.relu, [5])
    m.load_params(p)
    want = O.predict(O.RecurrentNet(p, network, 2), x, off, network, 1, td, None)
    got = nn.predict(m, x, off, 39, network, 0, 1, td, None, progress=False)
    assert np.abs(got - want).max() < 1e-3
    assert np.all(got[off[0]:off[1]] == 0) and np.all(got[off[2]:off[3]] == 0)  # shorter than the delay
    pinned = nn.empty_pinned((off[-1], 39))
    pinned[:] = 7.0  # a caller-owned buffer with stale contents: every row must still be written
    assert np.array_equal(nn.predict(m, x, off, 39, network, 0, 1, td, None, progress=False, out=pinned), got)
    empty = nn.predict(m, x[:0], np.zeros(1, np.int32), 39, network, 0, 1, td, None, progress=False)
    assert empty.shape == (0, 39)


@pytest.mark.parametrize("network,units,n_utt", [("lstm", 512, 300), ("blstm", 128, 300), ("lstm", 192, 300),
                                                 ("lstm", 512, 2700), ("blstm", 512, 2500)])
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_wide_kernel_128_slots_per_batch(nn, network, units, n_utt, prec):
    """The 128-slot "wide" LSTM kernel (utterances on the MMA's M axis, h streamed through a TMA ring) against the
    oracle and against the 32-slot kernel, on more than one batch with ragged lengths."""
    from nnacousticmodeling_b200 import recurrent_engine
    rng = np.random.default_rng(units)
    # 2500+ utterances: more batches than (CTA group, stream) lanes, so lanes run several items back to back
    lens = rng.integers(1, 60 if n_utt == 300 else 45, size=n_utt).tolist()
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network == "blstm"
    m, p = _lstm(nn, 77, network, 40, units, 2, 39, precision=prec, bidirectional=bid)
    wide = np.zeros((off[-1], 39), np.float32)
    recurrent_engine.forward_utterances(m, x, off, wide, 0, len(lens), timedelay=0, device=0, nb=128)
    narrow = np.zeros_like(wide)
    recurrent_engine.forward_utterances(m, x, off, narrow, 0, len(lens), timedelay=0, device=0, nb=32)
    assert np.abs(wide - narrow).max() < 2e-2  # same bf16 arithmetic, different summation order / gx rounding points
    pick = [0, 17, int(np.argmax(lens)), int(np.argmin(lens)), n_utt - 1, n_utt // 2]
    for u in pick:
        xs = x[off[u]:off[u + 1]]
        if bid:
            want = O.log_softmax(O.birnn_forward_utterance(p, "lstm", 2, xs))
        else:
            want = O.log_softmax(O.rnn_forward_utterance(p, "lstm", 2, xs))
        assert np.abs(wide[off[u]:off[u + 1]] - want).max() < 5e-2


@pytest.mark.parametrize("network,units,n_utt", [("gru", 512, 300), ("mgrurelu", 512, 300), ("mgrurelur", 128, 300),
                                                 ("bgru", 192, 300), ("gru", 512, 2700), ("mgrurelu", 256, 2600)])
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_wide_gru_kernel_128_slots_per_batch(nn, network, units, n_utt, prec):
    """The 128-slot GRU-family kernel (gate-blocked rows, two exchanges per step with the reset gate) against the
    32-slot kernel on the whole set and against the oracle on a few utterances."""
    from nnacousticmodeling_b200 import recurrent_engine
    rng = np.random.default_rng(units + n_utt)
    lens = rng.integers(1, 60 if n_utt == 300 else 45, size=n_utt).tolist()
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network == "bgru"
    m, p = _gru(nn, 91, network, 40, units, 2, 39, precision=prec, bidirectional=bid)
    wide = np.zeros((off[-1], 39), np.float32)
    recurrent_engine.forward_utterances(m, x, off, wide, 0, len(lens), timedelay=0, device=0, nb=128)
    narrow = np.zeros_like(wide)
    recurrent_engine.forward_utterances(m, x, off, narrow, 0, len(lens), timedelay=0, device=0, nb=32)
    assert np.abs(wide - narrow).max() < 3e-2
    base = "gru" if bid else network
    for u in [0, 17, int(np.argmax(lens)), int(np.argmin(lens)), n_utt - 1, n_utt // 2]:
        xs = x[off[u]:off[u + 1]]
        if bid:
            want = O.log_softmax(O.birnn_forward_utterance(p, "gru", 2, xs))
        else:
            want = O.log_softmax(O.rnn_forward_utterance(p, base, 2, xs))
        assert np.abs(wide[off[u]:off[u + 1]] - want).max() < 5e-2


@pytest.mark.parametrize("network,units", [("lstm", 512), ("blstm", 512), ("gru", 512), ("bgru", 512), ("mgrurelu", 512)])
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_mixed_schedule_long_utterances_in_the_32_slot_kernel(nn, network, units, prec, monkeypatch):
    """MixedSchedule: the longest utterances run in the 32-slot kernel, the rest in the 128-slot kernel, as two
    concurrent launches over one packed row space; same outputs as the plain 32-slot schedule."""
    from nnacousticmodeling_b200 import recurrent_engine
    rng = np.random.default_rng(7 + units)
    lens = np.concatenate([rng.integers(150, 200, size=40), rng.integers(1, 70, size=700)])
    rng.shuffle(lens)
    off = _offsets(lens.tolist())
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network in ("blstm", "bgru")
    if "lstm" in network:
        m, p = _lstm(nn, 78, network, 40, units, 2, 39, precision=prec, bidirectional=bid)
    else:
        m, p = _gru(nn, 78, network, 40, units, 2, 39, precision=prec, bidirectional=bid)
    monkeypatch.setenv("NNAM_RNN_MIXED", "force")
    mixed = np.zeros((off[-1], 39), np.float32)
    recurrent_engine.forward_utterances(m, x, off, mixed, 0, len(lens), timedelay=3 if not bid else 0, device=0)
    plan = next(iter(m._plans.values()))
    assert any(isinstance(v[0], recurrent_engine.MixedSchedule) for v in plan._sched_cache.values())
    narrow = np.zeros_like(mixed)
    recurrent_engine.forward_utterances(m, x, off, narrow, 0, len(lens), timedelay=3 if not bid else 0, device=0, nb=32)
    assert np.abs(mixed - narrow).max() < 3e-2
    assert np.array_equal(mixed == 0, narrow == 0)  # the same rows stay unwritten (quirk Q4)


@pytest.mark.parametrize("network", ["lstm", "bgru"])
def test_two_phase_forward_equals_single_phase(nn, network, monkeypatch):
    """Large shards with a host output are computed in two subsets of utterances (shortest first) so that the first
    subset's device->host copy overlaps the second's computation: bit-identical to the single-phase result, including
    quirk Q4 rows and utterances shorter than the time delay."""
    from nnacousticmodeling_b200 import recurrent_engine
    rng = np.random.default_rng(17)
    lens = rng.integers(350, 750, size=300)
    lens[[5, 77, 123]] = [2, 1, 3]  # shorter than / equal to the delay
    rng.shuffle(lens)
    off = _offsets(lens.tolist())
    assert off[-1] >= 150000
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network == "bgru"
    td = 0 if bid else 3
    if network == "lstm":
        m, _ = _lstm(nn, 5, network, 40, 64, 2, 39, precision="bf16")
    else:
        m, _ = _gru(nn, 5, network, 40, 64, 2, 39, precision="bf16", bidirectional=True)
    assert len(recurrent_engine._phase_split(lens)) == 2
    two = np.full((off[-1], 39), 7.0, np.float32)
    recurrent_engine.forward_utterances(m, x, off, two, 0, len(lens), timedelay=td, device=0)
    monkeypatch.setenv("NNAM_RNN_PHASES", "1")
    assert len(recurrent_engine._phase_split(lens)) == 1
    one = np.full((off[-1], 39), 7.0, np.float32)
    recurrent_engine.forward_utterances(m, x, off, one, 0, len(lens), timedelay=td, device=0)
    assert np.array_equal(two, one)
    if td:
        for u in (int(np.argmin(lens)), 0):
            assert np.all(two[off[u + 1] - min(td, lens[u]):off[u + 1]] == 0)


@pytest.mark.parametrize("network,nb", [("lstm", 32), ("lstm", 128), ("gru", 32), ("gru", 128), ("blstm", 128),
                                        ("peepholelstm", 32)])
def test_exchange_buffer_contents_on_entry_are_irrelevant(nn, network, nb):
    """include/nnam_b200.h: the kernels never read an exchange slot row they have not written in the same launch.  The
    buffer is reused across layers, models and schedules without being cleared, so it is poisoned with NaNs here: any
    read-before-write would surface as a NaN in the output."""
    from nnacousticmodeling_b200 import engine, recurrent_engine
    rng = np.random.default_rng(23)
    lens = rng.integers(1, 50, size=200).tolist()
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network == "blstm"
    if "lstm" in network:
        m, _ = _lstm(nn, 31, network, 40, 128, 2, 39, precision="fp16", bidirectional=bid)
    else:
        m, _ = _gru(nn, 31, network, 40, 128, 2, 39, precision="fp16")
    clean = np.zeros((off[-1], 39), np.float32)
    kw = dict(timedelay=0, device=0) if network == "peepholelstm" else dict(timedelay=0, device=0, nb=nb)
    recurrent_engine.forward_utterances(m, x, off, clean, 0, len(lens), **kw)
    plan = engine.get_plan(m, 0)
    poisoned = 0
    for name, buf in plan.ws.buf.items():
        if "xchg" in name:
            buf.fill_(float("nan"))
            poisoned += 1
    assert poisoned >= 1
    again = np.zeros_like(clean)
    recurrent_engine.forward_utterances(m, x, off, again, 0, len(lens), **kw)
    assert np.isfinite(again).all() and np.array_equal(again, clean)


def test_recurrent_compact_transfer(nn):
    """Recurrent path with transfer="f16" (the default of the 16-bit modes): compact rows through the pinned staging area, widened on the
    host per phase; quirk Q4 rows stay exactly 0, everything else within the format's bound of the float32 transfer."""
    rng = np.random.default_rng(41)
    lens = rng.integers(2, 70, size=260).tolist()
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    m, _ = _lstm(nn, 9, "lstm", 40, 128, 2, 1909, precision="fp16")
    f32 = nn.predict(m, x, off, 1909, "lstm", 0, 1, 3, None, progress=False, transfer="f32")
    f16 = nn.predict(m, x, off, 1909, "lstm", 0, 1, 3, None, progress=False, transfer="f16")
    from nnacousticmodeling_b200 import engine  # default: compact when the process has the threads to widen with
    auto = f16 if engine.default_host_threads() >= engine.MIN_WIDEN_THREADS else f32
    assert np.array_equal(nn.predict(m, x, off, 1909, "lstm", 0, 1, 3, None, progress=False), auto)
    m.precision = "fp32"  # the fp32-accurate mode keeps float32 rows
    a = nn.predict(m, x, off, 1909, "lstm", 0, 1, 3, None, progress=False)
    assert np.array_equal(a, nn.predict(m, x, off, 1909, "lstm", 0, 1, 3, None, progress=False, transfer="f32"))
    m.precision = "fp16"
    assert np.array_equal(f16 == 0, f32 == 0)
    dist = f32.max(axis=1, keepdims=True) - f32
    assert np.all(np.abs(f16 - f32) <= 2.0 ** -11 * dist + 2e-5)
    live = np.abs(f32).sum(axis=1) > 0
    assert np.array_equal(f16[live].argmax(axis=1), f32[live].argmax(axis=1))


@pytest.mark.parametrize("network,nb", [("lstm", 32), ("lstm", 64), ("lstm", 128), ("blstm", None), ("gru", 32),
                                        ("gru", 128), ("bgru", None), ("peepholelstm", None)])
def test_repeated_runs_are_bit_identical(nn, network, nb, monkeypatch):
    """Race proxy (compute-sanitizer is closed on this GPU pool, profiles/r02_sanitizer.md): every K3 kernel family --
    counter-based exchange, cluster + multicast exchange, 128-slot two-stream kernels, the mixed two-launch schedule
    (nb=None lets the cost model pick it; forced here) -- must give bit-identical outputs on repeated runs of the same
    ragged workload.  A missing barrier / fence in the cross-CTA exchange shows up as run-to-run differences."""
    from nnacousticmodeling_b200 import recurrent_engine
    rng = np.random.default_rng(97)
    lens = np.concatenate([rng.integers(120, 160, size=40), rng.integers(1, 60, size=500)])
    rng.shuffle(lens)
    off = _offsets(lens.tolist())
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network in ("blstm", "bgru")
    if "lstm" in network:
        m, _ = _lstm(nn, 55, network, 40, 512 if network != "peepholelstm" else 128, 2, 39, precision="fp16",
                     bidirectional=bid)
    else:
        m, _ = _gru(nn, 55, network, 40, 512, 2, 39, precision="fp16", bidirectional=bid)
    if nb is None and network != "peepholelstm":
        monkeypatch.setenv("NNAM_RNN_MIXED", "force")
    kw = {} if nb is None else {"nb": nb}
    outs = []
    for _ in range(6):
        o = np.zeros((off[-1], 39), np.float32)
        recurrent_engine.forward_utterances(m, x, off, o, 0, len(lens), timedelay=0, device=0, **kw)
        outs.append(o)
    assert np.isfinite(outs[0]).all()
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])


@pytest.mark.parametrize("network,precision", [("lstm", "fp16"), ("lstm", "fp32"), ("blstm", "fp16"), ("gru", "bf16")])
def test_fused_output_layer_matches_unfused_recurrent(nn, monkeypatch, network, precision):
    """Recurrent path: the fused output layer + head (default for 512..2048 classes) scatters the packed time-major rows to
    their frames, drops the delay rows and zero-fills the quirk-Q4 rows exactly like nnam_head_scatter."""
    rng = np.random.default_rng(43)
    lens = rng.integers(2, 90, size=300).tolist()
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    bid = network == "blstm"
    m, _ = _lstm(nn, 11, network, 40, 128, 2, 1909, precision=precision, bidirectional=bid)
    td = 0 if bid else 3
    res = {}
    for env in ("0", "1"):
        monkeypatch.setenv("NNAM_FUSED_HEAD", env)
        res[env] = (nn.predict(m, x, off, 1909, network, 0, 1, td, None, progress=False, transfer="f32"),
                    nn.predict(m, x, off, 1909, network, 0, 1, td, None, progress=False, transfer="f16"))
    u, f = res["0"][0], res["1"][0]
    assert np.array_equal(u == 0, f == 0)
    assert np.abs(u - f).max() < 3e-5
    u, f = res["0"][1], res["1"][1]
    assert np.array_equal(u == 0, f == 0)
    assert np.all(np.abs(u - f) <= 2.0 ** -10 * (u.max(axis=1, keepdims=True) - u) + 3e-5)
