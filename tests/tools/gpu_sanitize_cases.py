#!/usr/bin/env python3
"""Small instances of every kernel family for compute-sanitizer (racecheck / synccheck / memcheck / initcheck):

  compute-sanitizer --tool racecheck python tests/tools/gpu_sanitize_cases.py [case ...]

Cases: k1 (splice stream + tile kernels, gather, convert), k2 (1-CTA and cta_group::2 GEMM, every output kind / pass
mode), k3_64 (32/64-slot LSTM, GRU, peephole), k3_128 (128-slot LSTM / GRU, two streams per group), k3_mixed (two
concurrent cooperative launches), k4 (head variants).  Every case checks its result against the oracle, so a run that
the tool slows down 50x is still a correctness run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import nnacousticmodeling_b200 as nn  # noqa: E402
from nnacousticmodeling_b200 import ops, recurrent_engine  # noqa: E402
from oracle import nnam_oracle as O  # noqa: E402

DEV = torch.device("cuda:0")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def case_k1():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3000, 40)).astype(np.float32)
    iv = rng.standard_normal((3000, 100)).astype(np.float32)
    ft = O.load_kaldi_feature_transform(os.path.join(ROOT, "tests", "golden", "final.feature_transform"))
    want = np.concatenate((O.apply_kaldi_feature_transform(O.splicing(x, range(-5, 6)), ft), iv), axis=1)
    for kind in (ops.OUT_F32, ops.OUT_BF16, ops.OUT_F16, ops.OUT_BF16_SPLIT):
        hi, lo = ops.splice_transform(t(x), 3000, 5, t(ft["addShift"]), t(ft["rescale"]), t(iv), out_kind=kind)
        got = hi.float() + (lo.float() if lo is not None else 0)
        assert np.abs(got.cpu().numpy()[:, :540] - want).max() < (1e-6 if kind == ops.OUT_F32 else 3e-2)
    rmap = t(rng.integers(0, 3000, 999).astype(np.int32))
    ops.gather_transform(t(x), rmap, None, None, t(iv), out_kind=ops.OUT_F16)
    ops.convert_f32(t(x), ops.OUT_BF16_SPLIT)


def case_k2():
    rng = np.random.default_rng(1)
    for m, n, k in ((300, 520, 136), (4200, 1909, 264)):  # single-CTA kernel; CTA-pair kernel (M >= 4096, K >= 384 no: 264 -> 1-CTA)
        for mk in (k, 520):
            a = rng.standard_normal((m, mk)).astype(np.float32)
            w = (rng.standard_normal((n, mk)) / np.sqrt(mk)).astype(np.float32)
            b = rng.standard_normal(n).astype(np.float32)
            ref = a.astype(np.float64) @ w.astype(np.float64).T + b
            a_hi, a_lo = ops.convert_f32(t(a), ops.OUT_BF16_SPLIT)
            w_hi, w_lo = ops.convert_f32(t(w), ops.OUT_BF16_SPLIT)
            for ns, tol in ((1, 3e-2), (2, 2e-2), (3, 3e-4), (4, 2e-2)):
                got, _ = ops.linear_bias_act(a_hi, a_lo, w_hi, w_lo, t(b), m, n, mk, out_kind=ops.OUT_F32, nsplit=ns)
                assert np.abs(got.cpu().numpy()[:, :n] - ref).max() < tol * max(1.0, np.abs(ref).max()), (m, n, mk, ns)
            ops.linear_bias_act(a_hi, a_lo, w_hi, w_lo, t(b), m, n, mk, act="relu", out_kind=ops.OUT_BF16_SPLIT, nsplit=3)
            h_a, _ = ops.convert_f32(t(a), ops.OUT_F16)
            h_w, _ = ops.convert_f32(t(w), ops.OUT_F16)
            ops.linear_bias_act(h_a, None, h_w, None, t(b), m, n, mk, act="tanh", out_kind=ops.OUT_F16, elem=ops.ELEM_F16)
    torch.cuda.synchronize()


def _rnn(network, units, n_utt, max_len, nb, prec, bid=False, mixed=False):
    rng = np.random.default_rng(units + n_utt)
    lens = rng.integers(1, max_len, size=n_utt)
    if mixed:
        lens[:40] = rng.integers(3 * max_len, 4 * max_len, size=40)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    base = {"blstm": "lstm", "bgru": "gru"}.get(network, network)
    p = O.init_recurrent(np.random.default_rng(5), base, 40, units, 2, 39, bidirectional=bid, bias_scale=0.1)
    m = nn.get_nn(network, 2, [units], 39, nn.F.relu, [5])
    m.load_params(p)
    m.precision = prec
    out = np.zeros((off[-1], 39), np.float32)
    kw = {} if nb is None else {"nb": nb}
    recurrent_engine.forward_utterances(m, x, off, out, 0, n_utt, timedelay=0, device=0, **kw)
    for u in (0, n_utt // 2, int(np.argmax(lens))):
        xs = x[off[u]:off[u + 1]]
        want = O.log_softmax(O.birnn_forward_utterance(p, base, 2, xs) if bid else O.rnn_forward_utterance(p, base, 2, xs))
        assert np.abs(out[off[u]:off[u + 1]] - want).max() < (1e-3 if prec == "fp32" else 5e-2), (network, nb, prec)


def case_k3_64():
    _rnn("lstm", 128, 70, 9, 32, "fp16")
    _rnn("lstm", 128, 70, 9, 64, "bf16")
    _rnn("lstm", 64, 40, 7, 16, "fp32")
    _rnn("gru", 128, 70, 9, 32, "fp16")
    _rnn("mgrurelu", 128, 70, 9, 64, "bf16")
    _rnn("bgru", 64, 40, 7, 32, "fp32", bid=True)
    _rnn("peepholelstm", 128, 40, 7, None, "fp16")


def case_k3_128():
    _rnn("lstm", 128, 400, 8, 128, "fp16")
    _rnn("blstm", 128, 300, 8, 128, "bf16", bid=True)
    _rnn("gru", 128, 400, 8, 128, "fp16")
    _rnn("mgrurelu", 128, 300, 8, 128, "bf16")


def case_k3_mixed():
    os.environ["NNAM_RNN_MIXED"] = "force"
    _rnn("lstm", 128, 500, 8, None, "fp16", mixed=True)
    _rnn("bgru", 128, 400, 8, None, "fp16", bid=True, mixed=True)
    os.environ.pop("NNAM_RNN_MIXED")


def case_k4():
    rng = np.random.default_rng(2)
    ys = [rng.standard_normal((777, 1909)).astype(np.float32) for _ in range(3)]
    yd = []
    for a in ys:
        buf = torch.zeros(777, 1920, device=DEV)
        buf[:, :1909] = t(a)
        yd.append(buf)
    ap = rng.standard_normal(1909).astype(np.float32)
    got = ops.head(yd[0], 1909, prior=t(ap)).cpu().numpy()
    assert np.abs(got - O.head(ys[0], ap[None, :])).max() < 1e-4
    ops.head(yd, 1909, weights=[0.5, 0.25, 0.25])
    ops.head(yd, 1909, pre_normalize=True)
    rmap = t(rng.permutation(777).astype(np.int32))
    ops.head(yd[1], 1909, out=torch.zeros(777, 1909, device=DEV), out_row_map=rmap)
    torch.cuda.synchronize()


CASES = {"k1": case_k1, "k2": case_k2, "k3_64": case_k3_64, "k3_128": case_k3_128, "k3_mixed": case_k3_mixed, "k4": case_k4}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
        torch.cuda.synchronize()
        print("case", name, "ok", flush=True)
