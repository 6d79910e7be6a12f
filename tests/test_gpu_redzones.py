"""Out-of-bounds write checks of our own (no external memory checker is available on the GPU pool): every kernel
writes into a view carved out of a guard-filled buffer, and the guard bytes plus the inputs must be unchanged
afterwards.  Kernel level through the C ABI (ops.*) with ragged shapes, and engine level with the workspace's
NNAM_REDZONE mode on for every network family.  -m gpu."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nnam_oracle as O  # noqa: E402

GUARD = 0xA5
PAD = 8192  # guard bytes either side


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from nnacousticmodeling_b200 import ops as _ops
    return _ops


class Guarded:
    """(rows, ld) tensor of `dtype` with PAD guard bytes before and after it."""

    def __init__(self, rows, ld, dtype, dev):
        self.nbytes = rows * ld * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((2 * PAD + self.nbytes,), GUARD, dtype=torch.uint8, device=dev)
        self.t = self.raw[PAD:PAD + self.nbytes].view(dtype).view(rows, ld)

    def assert_intact(self, what):
        torch.cuda.synchronize()
        front = self.raw[:PAD].cpu().numpy()
        back = self.raw[PAD + self.nbytes:].cpu().numpy()
        assert np.all(front == GUARD), f"{what}: wrote before the buffer at byte {-PAD + int(np.argmax(front != GUARD))}"
        assert np.all(back == GUARD), f"{what}: wrote past the buffer at byte +{int(np.argmax(back != GUARD))}"


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("n,dim,splice,ivd,f0,f1", [(1000, 40, 5, 100, 0, 1000), (333, 40, 5, 0, 7, 330),
                                                   (70, 13, 3, 5, 0, 70), (129, 40, 0, 100, 1, 128),
                                                   (1, 40, 5, 100, 0, 1), (257, 40, 5, 3, 250, 257)])
@pytest.mark.parametrize("kind", ["f32", "bf16", "split"])
def test_splice_stays_inside_its_output(ops, dev, n, dim, splice, ivd, f0, f1, kind):
    rng = np.random.default_rng(n + dim)
    x = _t(rng.standard_normal((n, dim)).astype(np.float32), dev)
    iv = _t(rng.standard_normal((f1 - f0, ivd)).astype(np.float32), dev) if ivd else None
    cols = (2 * splice + 1) * dim
    add, mul = _t(rng.standard_normal(cols).astype(np.float32), dev), _t(rng.standard_normal(cols).astype(np.float32), dev)
    ok = {"f32": ops.OUT_F32, "bf16": ops.OUT_BF16, "split": ops.OUT_BF16_SPLIT}[kind]
    ldo = cols + ivd if kind == "f32" else ops.round_up(cols + ivd, 8)
    dt = torch.float32 if kind == "f32" else torch.bfloat16
    hi = Guarded(f1 - f0, ldo, dt, dev)
    lo = Guarded(f1 - f0, ldo, dt, dev) if kind == "split" else None
    x0 = x.clone()
    ops.splice_transform(x, n, splice, add, mul, iv, f0=f0, f1=f1, out_kind=ok, ldo=ldo,
                         out=(hi.t, lo.t if lo else None))
    hi.assert_intact("splice hi")
    if lo:
        lo.assert_intact("splice lo")
    assert torch.equal(x, x0)
    want = O.apply_kaldi_feature_transform(O.splicing(x0.cpu().numpy(), range(-splice, splice + 1)),
                                           {"addShift": add.cpu().numpy(), "rescale": mul.cpu().numpy()})[f0:f1]
    got = hi.t[:, :cols].float().cpu().numpy()
    assert np.abs(got - want).max() <= (0 if kind == "f32" else 2 ** -7 * np.abs(want).max())


@pytest.mark.parametrize("m,n,k,ldo", [(1, 16, 8, 16), (77, 40, 40, 48), (300, 1024, 544, 1024), (1000, 1909, 440, 1920),
                                       (129, 2048, 2048, 2048), (4097, 256, 1024, 256), (127, 1909, 512, 1920)])
@pytest.mark.parametrize("kind", ["f32", "bf16", "split"])
def test_linear_stays_inside_its_output(ops, dev, m, n, k, ldo, kind):
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    split = kind == "split"
    a_hi, a_lo = ops.convert_f32(_t(a, dev), ops.OUT_BF16_SPLIT if split else ops.OUT_BF16)
    w_hi, w_lo = ops.convert_f32(_t(w, dev), ops.OUT_BF16_SPLIT if split else ops.OUT_BF16)
    ok = {"f32": ops.OUT_F32, "bf16": ops.OUT_BF16, "split": ops.OUT_BF16_SPLIT}[kind]
    dt = torch.float32 if kind == "f32" else torch.bfloat16
    hi = Guarded(m, ldo, dt, dev)
    lo = Guarded(m, ldo, dt, dev) if split else None
    a_copy, w_copy = a_hi.clone(), w_hi.clone()
    ops.linear_bias_act(a_hi, a_lo, w_hi, w_lo, _t(b, dev), m, n, k, "relu", ok, 3 if split else 1,
                        out=(hi.t, lo.t if lo else None), ldo=ldo)
    hi.assert_intact(f"gemm {m}x{n}x{k} hi")
    if lo:
        lo.assert_intact(f"gemm {m}x{n}x{k} lo")
    assert torch.equal(a_hi, a_copy) and torch.equal(w_hi, w_copy)
    want = np.maximum(a_hi.float().cpu().numpy() @ w_hi.float().cpu().numpy()[:, :k].T + b, 0) if not split else \
        np.maximum(a.astype(np.float64) @ w.astype(np.float64).T + b, 0)
    got = hi.t[:, :n].float().cpu().numpy()
    if split:
        got = got + lo.t[:, :n].float().cpu().numpy()
    tol = 1e-3 if split else (2e-2 if kind != "f32" else 2e-3)
    assert np.abs(got - want).max() < tol * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("c,rows,ld,k", [(39, 77, 48, 1), (1909, 300, 1920, 1), (1909, 33, 1920, 3), (7, 1, 7, 2),
                                         (2048, 9, 2048, 1), (1000, 17, 1000, 1)])
@pytest.mark.parametrize("scatter", [False, True])
def test_head_stays_inside_its_output(ops, dev, c, rows, ld, k, scatter):
    rng = np.random.default_rng(c + rows)
    logits = [_t(rng.standard_normal((rows, ld)).astype(np.float32), dev) for _ in range(k)]
    prior = _t(rng.standard_normal(c).astype(np.float32), dev)
    n_out_rows = rows + 3 if scatter else rows
    out = Guarded(n_out_rows, c, torch.float32, dev)
    out.t.fill_(7.0)
    row_map = None
    if scatter:  # a permutation with two dropped rows and the three spare output rows zero-filled through rows 0..2
        perm = rng.permutation(rows).astype(np.int32)
        row_map_np = perm.copy()
        dropped = []
        if rows > 4:
            dropped = [int(perm[3]), int(perm[4])]
            row_map_np[3] = -2 - rows          # zero-fill output row `rows`
            row_map_np[4] = -1                 # drop
        row_map = _t(row_map_np, dev)
    ops.head(logits, c, rows=rows, weights=[1.0 / k] * k if k > 1 else None, prior=prior, prior_scale=0.5, out=out.t,
             out_row_map=row_map)
    out.assert_intact("head")
    z = sum(l.cpu().numpy()[:, :c].astype(np.float64) for l in logits) / k - 0.5 * prior.cpu().numpy()
    want = z - np.log(np.exp(z - z.max(axis=1, keepdims=True)).sum(axis=1, keepdims=True)) - z.max(axis=1, keepdims=True)
    got = out.t.cpu().numpy()
    if not scatter:
        assert np.abs(got - want).max() < 1e-4
    else:
        for r in range(rows):
            q = int(row_map_np[r])
            if q >= 0:
                assert np.abs(got[q] - want[r]).max() < 1e-4
        if rows > 4:
            assert np.all(got[rows] == 0)               # zero-filled
            for q in dropped:
                assert np.all(got[q] == 7.0)            # nobody wrote the rows whose source was dropped / redirected
            assert np.all(got[rows + 1:] == 7.0)


@pytest.mark.parametrize("n_src,dim,ivd,n_rows", [(100, 40, 100, 257), (50, 40, 0, 1), (64, 13, 5, 129)])
@pytest.mark.parametrize("kind", ["bf16", "split"])
def test_gather_stays_inside_its_output(ops, dev, n_src, dim, ivd, n_rows, kind):
    rng = np.random.default_rng(n_src + n_rows)
    x = _t(rng.standard_normal((n_src, dim)).astype(np.float32), dev)
    iv = _t(rng.standard_normal((n_src, ivd)).astype(np.float32), dev) if ivd else None
    rm_np = rng.integers(-1, n_src, n_rows).astype(np.int32)  # -1 = zero row
    rm = _t(rm_np, dev)
    ldo = ops.round_up(dim + ivd, 8)
    ok = ops.OUT_BF16 if kind == "bf16" else ops.OUT_BF16_SPLIT
    hi = Guarded(n_rows, ldo, torch.bfloat16, dev)
    lo = Guarded(n_rows, ldo, torch.bfloat16, dev) if kind == "split" else None
    ops.gather_transform(x, rm, None, None, iv, ok, ldo, out=(hi.t, lo.t if lo else None))
    hi.assert_intact("gather hi")
    if lo:
        lo.assert_intact("gather lo")
    got = hi.t[:, :dim].float().cpu().numpy()
    want = np.where(rm_np[:, None] >= 0, x.cpu().numpy()[np.maximum(rm_np, 0)], 0)
    assert np.abs(got - want).max() <= 2 ** -7 * np.abs(want).max()


# ------------------------------------------------------------------------------------------ engine level
def _fresh(nn, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    monkeypatch.setenv("NNAM_REDZONE", "16384")


def _check_plans(m):
    n = 0
    for plan in m._plans.values():
        n += plan.ws.check_redzones()
    assert n > 0, "no guarded workspace buffers were checked (NNAM_REDZONE not honoured?)"


def _offsets(lens):
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_ff_and_tdnn_engines_keep_inside_their_workspaces(monkeypatch, precision):
    import nnacousticmodeling_b200 as nn
    _fresh(nn, monkeypatch)
    x, _, iv = O.synth_set(31, 9, ivec_dim=100)
    x, iv = x[:2777], iv[:2777]
    m = nn.get_nn("ff", 3, [512], 1909, "relu", [5])
    m.init_params(540, np.random.default_rng(1))
    m.precision = precision
    got = nn.predict(m, x, None, 1909, "ff", 0, 11, 0, None, progress=False, ivectors=iv)
    assert np.isfinite(got).all()
    _check_plans(m)
    t = nn.get_nn("tdnn", 4, [128, 128, 128, 128], 39, "relu", [5, 5, 5, 5])
    t.init_params(40 * 17, np.random.default_rng(2))
    t.precision = precision
    got = nn.predict(t, x, None, 39, "tdnn", 0, 17, 0, None, progress=False)
    assert np.isfinite(got).all()
    _check_plans(t)


@pytest.mark.parametrize("network,units,n_utt", [("lstm", 512, 9), ("lstm", 192, 70), ("blstm", 128, 40),
                                                 ("lstm", 512, 140), ("gru", 128, 40), ("mgrurelur", 64, 5),
                                                 ("bgru", 128, 37), ("peepholelstm", 128, 20), ("peepholelstm", 512, 6)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_recurrent_engines_keep_inside_their_workspaces(monkeypatch, network, units, n_utt, precision):
    import nnacousticmodeling_b200 as nn
    _fresh(nn, monkeypatch)
    rng = np.random.default_rng(units + n_utt)
    lens = rng.integers(1, 60, n_utt)
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    base = {"blstm": "lstm", "bgru": "gru"}.get(network, network)
    p = O.init_recurrent(np.random.default_rng(5), base, 40, units, 2, 39, bidirectional=network.startswith("b"))
    m = nn.get_nn(network, 2, [units], 39, nn.F.relu, [5])
    m.load_params(p)
    m.precision = precision
    bid = network.startswith("b")
    td = 0 if bid else 2
    got = nn.predict(m, x, off, 39, network, 0, 1, td, None, progress=False)
    assert np.isfinite(got).all()
    _check_plans(m)
    if bid:
        want = np.concatenate([O.log_softmax(O.birnn_forward_utterance(p, base, 2, x[off[u]:off[u + 1]]))
                               for u in range(n_utt)])
    else:
        want = O.predict(O.RecurrentNet(p, network, 2), x, off, network, 1, td, None)
    assert np.abs(got - want).max() < (1e-3 if precision == "fp32" else 5e-2)


def test_mixed_schedule_keeps_inside_its_workspaces(monkeypatch):
    """Two concurrent recurrence launches over one packed row space (MixedSchedule), guard bytes around every buffer."""
    import nnacousticmodeling_b200 as nn
    from nnacousticmodeling_b200 import recurrent_engine
    _fresh(nn, monkeypatch)
    monkeypatch.setenv("NNAM_RNN_MIXED", "force")
    rng = np.random.default_rng(21)
    lens = np.concatenate([rng.integers(90, 130, size=40), rng.integers(1, 50, size=300)])
    off = _offsets(lens)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    p = O.init_recurrent(np.random.default_rng(6), "lstm", 40, 512, 2, 39)
    m = nn.get_nn("lstm", 2, [512], 39, nn.F.relu, [5])
    m.load_params(p)
    m.precision = "bf16"
    got = nn.predict(m, x, off, 39, "lstm", 0, 1, 2, None, progress=False)
    plan = next(iter(m._plans.values()))
    assert any(isinstance(v[0], recurrent_engine.MixedSchedule) for v in plan._sched_cache.values())
    _check_plans(m)
    want = O.predict(O.RecurrentNet(p, "lstm", 2), x, off, "lstm", 1, 2, None)
    assert np.abs(got - want).max() < 5e-2
