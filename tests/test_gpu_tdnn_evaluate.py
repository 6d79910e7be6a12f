"""TDNN (conv-as-GEMM), NNWithRPL ensembles, evaluateModelTestTri / .lab and the dev-mode CLI against the oracle.
-m gpu."""
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


def _offsets(lens):
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)


def _tdnn(nn, seed, units, ksize, n_out, act="relu"):
    m = nn.get_nn("tdnn", len(ksize), units, n_out, act, ksize)
    win = sum(ksize) - len(ksize) + 1
    m.init_params(40 * win, np.random.default_rng(seed))
    p = dict(m.params)
    rng = np.random.default_rng(seed + 1)
    for k in p:
        if k.endswith("/b"):
            p[k] = (0.1 * rng.standard_normal(p[k].shape)).astype(np.float32)
    m.load_params(p)
    return m, p


@pytest.mark.parametrize("units,ksize,act", [([64, 64, 64, 64], [5, 5, 5, 5], "relu"), ([128, 96], [3, 7], "tanh"),
                                             ([512, 512, 512, 512], [5, 5, 5, 5], "sigmoid")])
def test_predict_tdnn(nn, golden_dir, units, ksize, act):
    """chainer_networks.py:24-42 incl. the scrambled reshape (quirk Q3), through predict() with the tiled transform."""
    n_out = 1909 if units[0] == 512 else 39
    m, p = _tdnn(nn, 21, units, ksize, n_out, act)
    win = sum(ksize) - len(ksize) + 1
    splice = win // 2
    x, _, _ = O.synth_set(22, 6)
    x = x[:1200]
    ft_full = nn.loadKaldiFeatureTransform(os.path.join(golden_dir, "final.feature_transform"))
    ft = nn.adapt_transform(ft_full, "tdnn", splice, False)
    oft = O.select_transform_for_network(O.load_kaldi_feature_transform(
        os.path.join(golden_dir, "final.feature_transform")), "tdnn", splice=splice)
    want = O.predict(lambda v: O.tdnn_forward(p, v, ksize, act), x, None, "tdnn", win, 0, oft)
    got = nn.predict(m, x, None, n_out, "tdnn", 0, win, 0, ft, progress=False)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-3
    m.precision = "bf16"
    got16 = nn.predict(m, x, None, n_out, "tdnn", 0, win, 0, ft, progress=False)
    assert np.abs(got16 - want).max() < 5e-2
    # model(x) surface on pre-spliced rows
    m.precision = "fp32"
    feats = O.apply_kaldi_feature_transform(O.splicing(x[:300], range(-splice, splice + 1)), oft)
    assert np.abs(m(feats) - O.tdnn_forward(p, feats, ksize, act)).max() < 1e-3


def _mlp(nn, seed, in_dim, units, layers, n_out):
    p = O.init_mlp(np.random.default_rng(seed), in_dim, units, layers, n_out)
    rng = np.random.default_rng(seed + 1)
    for k in p:
        if k.endswith("/b"):
            p[k] = (0.1 * rng.standard_normal(p[k].shape)).astype(np.float32)
    m = nn.get_nn("ff", layers, [units], n_out, "relu", [5])
    m.load_params(p)
    return m, p


def _rpl(nn, seed, n_out):
    rng = np.random.default_rng(seed)
    rp = {"W": (0.2 * rng.standard_normal((1, n_out))).astype(np.float32),
          "b": (0.1 * rng.standard_normal((1, n_out))).astype(np.float32),
          "lb": (-3.0 + 0.5 * rng.standard_normal((1, n_out))).astype(np.float32)}
    r = nn.RPL4(n_out)
    r.load_params(rp)
    return r, rp


@pytest.mark.parametrize("with_master,n_folds,with_rpl", [(True, 0, False), (True, 3, True), (False, 4, False),
                                                          (False, 2, True)])
def test_nn_with_rpl_ff_evaluate_forward(nn, golden_dir, with_master, n_folds, with_rpl):
    """evaluate.py:35-51 + RPL.py:68-74 + `y - ap; y - logsum(y)` (evaluateModelForTest.py:110-112)."""
    n_out = 1909
    ap = 0.8 * np.load(os.path.join(golden_dir, "log_ap_Kaldi1909.npy")).astype(np.float32)
    lens = [120, 77, 201, 33]
    off = _offsets(lens)
    data = np.random.default_rng(31).standard_normal((off[-1], 440)).astype(np.float32)
    master = _mlp(nn, 40, 440, 256, 2, n_out) if with_master else None
    folds = [_mlp(nn, 50 + i, 440, 256, 2, n_out) for i in range(n_folds)]
    rpl = _rpl(nn, 60, n_out) if with_rpl else None
    model = nn.NNWithRPL(master[0] if master else None, [f[0] for f in folds], rpl[0] if rpl else None)
    fwd = lambda p: (lambda v: O.mlp_forward(p, v, 2))  # noqa: E731
    ref = lambda v: O.nn_with_rpl(fwd(master[1]) if master else None, [fwd(f[1]) for f in folds],  # noqa: E731
                                  (lambda h: O.rpl4(rpl[1], h)) if rpl else None, v)
    want = np.concatenate(O.evaluate_forward(ref, data, off, ap=ap.reshape(1, -1), rnn=False))
    got = nn.evaluate.evaluate_forward(model, data, off, ap=ap, GPUID=0, rnn=False)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-3
    # __call__ surface: averaged logits (after RPL4)
    assert np.abs(model(data[:100]) - ref(data[:100])).max() < 1e-3


def test_nn_with_rpl_recurrent_and_lab_files(nn, golden_dir, tmp_path):
    """Recurrent ensemble through evaluateModelTestTri: no time-delay compensation, .lab files byte-compatible."""
    n_out = 1909
    ap = np.load(os.path.join(golden_dir, "log_ap_Kaldi1909.npy")).astype(np.float32)
    lens = [40, 25, 61, 33, 18]
    off = _offsets(lens)
    data = np.random.default_rng(71).standard_normal((off[-1], 40)).astype(np.float32)
    nets = []
    for i in range(2):
        p = O.init_recurrent(np.random.default_rng(80 + i), "lstm", 40, 128, 2, n_out)
        m = nn.get_nn("lstm", 2, [128], n_out, nn.F.relu, [5])
        m.load_params(p)
        nets.append((m, p))
    model = nn.NNWithRPL(None, [m for m, _ in nets], None)

    class Ens:  # oracle-side stateful ensemble
        def __init__(self):
            self.nets = [O.RecurrentNet(p, "lstm", 2) for _, p in nets]

        def reset_state(self):
            for n in self.nets:
                n.reset_state()

        def __call__(self, v):
            return O.nn_with_rpl(None, self.nets, None, v)

    want = O.evaluate_forward(Ens(), data, off, ap=ap.reshape(1, -1), rnn=True)
    lst = tmp_path / "lists"
    lst.mkdir()
    names = [f"dr1/spk{i}/utt{i}" for i in range(len(lens))]
    (lst / "test.list").write_text("\n".join(names) + "\n")
    lab = tmp_path / "lab"
    res = nn.evaluateModelTestTri(model, data, off, 10, 1.0, ap=ap, GPUID=0, testOrDev="test", tmpDir=str(lab),
                                  uttlistdir=str(lst), recogdir=str(tmp_path), progress=False, rnn=True)
    assert res is None  # no PhoneRecog binary: decoding is out of scope
    scp = (lab / "test.scp").read_text().splitlines()
    assert len(scp) == len(lens)
    for i, name in enumerate(names):
        f = lab / (name + ".lab")
        assert scp[i] == str(f)
        raw = np.fromfile(str(f), dtype=np.uint32, count=2)
        assert tuple(raw) == (lens[i], n_out)
        assert os.path.getsize(str(f)) == 8 + 4 * lens[i] * n_out
        got = O.load_bin(str(f))
        assert np.abs(got - want[i]).max() < 1e-3
    # wrong utterance count: the reference prints and returns -1
    assert nn.evaluateModelTestTri(model, data, off[:-1], 10, 1.0, uttlistdir=str(lst), tmpDir=str(lab)) == -1


def test_predict_folds_cli_dev_mode_and_fold_mode(nn, golden_dir, tmp_path):
    """predict_folds.py:97-246: per-fold outputs and the dev-mode mean of fold log-softmax outputs (quirk Q5)."""
    import importlib
    P = importlib.import_module("nnacousticmodeling_b200.predict")  # the package attribute `predict` is the function
    n_folds, n_out = 3, 1909
    rng = np.random.default_rng(90)
    models = tmp_path / "models"
    fdata = tmp_path / "fold_data"
    fout = tmp_path / "fold_out"
    data_dir = tmp_path / "data"
    for d in (models, fdata, data_dir):
        d.mkdir()
    import shutil
    shutil.copy(os.path.join(golden_dir, "final.feature_transform"), str(data_dir / "final.feature_transform"))
    ps = []
    for k in range(n_folds):
        m, p = _mlp(nn, 100 + k, 440, 128, 2, n_out)
        nn.save_npz(str(models / f"fold_{k}.npz"), nn.Classifier(m))
        ps.append(p)
        np.save(str(fdata / f"data_{k}.npy"), rng.standard_normal((500 + 37 * k, 40)).astype(np.float32))
    dev = rng.standard_normal((800, 40)).astype(np.float32)
    np.save(str(data_dir / "data_dev.npy"), dev)
    common = ["--network", "ff", "--units", 128, "--layers", 2, "--splice", 5, "--ft", "final.feature_transform",
              "--tri", "--data-dir", str(data_dir), "--fold-model-dir", str(models), "--no-progress", "--gpu", 0]
    P.main(common + ["--fold-data-dir", str(fdata), "--fold-output-dir", str(fout)])
    oft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    for k in range(n_folds):
        got = np.load(str(fout / f"data_{k}.npy"))
        x = np.load(str(fdata / f"data_{k}.npy"))
        want = O.predict(lambda v: O.mlp_forward(ps[k], v, 2), x, None, "ff", 11, 0, oft)
        assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
        assert np.abs(got - want).max() < 1e-3
    P.main(common + ["--fold-output-dir", str(fout), "--fold-output-dev", "data_dev.npy"])
    got = np.load(str(fout / "data_dev.npy"))
    acc = 0
    for k in range(n_folds):
        acc = acc + O.predict(lambda v: O.mlp_forward(ps[k], v, 2), dev, None, "ff", 11, 0, oft)
    acc = acc / n_folds
    want = acc - O.logsum(acc, axis=1)
    assert np.abs(got - want).max() < 1e-3


def test_evaluate_cli_matches_reference_data_flow(nn, golden_dir, tmp_path):
    """evaluate.py:53-214: splice -> transform -> i-vector concat -> net -> minus ap_coef * log_ap -> log-softmax -> .lab."""
    import shutil
    n_out = 1909
    lens = [64, 80, 37]
    off = _offsets(lens)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((off[-1], 40)).astype(np.float32)
    iv = (0.5 * rng.standard_normal((off[-1], 20))).astype(np.float32)
    data_dir, recog, lists, ivd = (tmp_path / d for d in ("data", "recog", "lists", "iv"))
    for d in (data_dir, recog, lists, ivd):
        d.mkdir()
    shutil.copy(os.path.join(golden_dir, "final.feature_transform"), str(data_dir / "final.feature_transform"))
    shutil.copy(os.path.join(golden_dir, "log_ap_Kaldi1909.npy"), str(recog / "log_ap_Kaldi1909.npy"))
    np.save(str(data_dir / "data_test.npy"), x)
    np.save(str(lists / "offsets_test.npy"), off)
    np.save(str(ivd / "ivectors_test.npy"), iv)
    names = ["a/u0", "a/u1", "b/u2"]
    (lists / "test.list").write_text("\n".join(names) + "\n")
    m, p = _mlp(nn, 33, 460, 128, 2, n_out)
    nn.save_npz(str(tmp_path / "model.npz"), nn.Classifier(m))
    lab = tmp_path / "lab"
    per = nn.evaluate.main(["--network", "ff", "--model", str(tmp_path / "model.npz"), "--units", 128, "--layers", 2,
                            "--splice", 5, "--tri", "--data-dir", str(data_dir), "--offset-dir", str(lists),
                            "--ivector-dir", str(ivd), "--recog-dir", str(recog), "--utt-list-dir", str(lists),
                            "--ap-coef", 0.7, "--no-progress", "--tmp-dir", str(lab)])
    assert per is None  # decoder not present
    oft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    feats = np.concatenate((O.apply_kaldi_feature_transform(O.splicing(x, range(-5, 6)), oft), iv), axis=1)
    ap = (np.float32(0.7) * np.load(os.path.join(golden_dir, "log_ap_Kaldi1909.npy")).astype(np.float32)).reshape(1, -1)
    want = O.evaluate_forward(lambda v: O.mlp_forward(p, v, 2), feats, off, ap=ap, rnn=False)
    for i, name in enumerate(names):
        got = O.load_bin(str(lab / (name + ".lab")))
        assert got.shape == (lens[i], n_out)
        assert np.abs(got - want[i]).max() < 1e-3
