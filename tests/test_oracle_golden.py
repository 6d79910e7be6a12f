"""The oracle against vectors produced by the reference's own NumPy helpers
(oracle/make_golden.py ran kw_utils.py / kw_nn_utils.py / orcus_util.py unmodified)."""
import os

import numpy as np

from oracle import nnam_oracle as O


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_feature_transform_parser(golden_dir):
    ft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    ref = _g(golden_dir, "feature_transform_parsed.npz")
    assert ft["shape"] == list(ref["shape"]) == [440, 40]
    assert ft["shifts"] == list(ref["shifts"]) == list(range(-5, 6))
    assert np.array_equal(ft["addShift"], ref["addShift"])
    assert np.array_equal(ft["rescale"], ref["rescale"])
    # the 11 per-shift blocks differ, so "select the middle block" is observable
    blocks = ft["addShift"].reshape(11, 40)
    assert np.abs(blocks - blocks[5]).max() > 1e-3


def test_splice_matches_both_reference_implementations(golden_dir):
    g = _g(golden_dir, "splice.npz")
    x = g["x"]
    assert np.array_equal(O.splicing(x, range(-5, 6)), g["splice11"])
    assert np.array_equal(O.prepare_batch(x, np.arange(len(x)), 11), g["splice11"])
    assert np.array_equal(O.prepare_batch(x, g["idx"], 11), g["prepare_sub"])
    assert np.array_equal(O.splicing(g["x_small"], range(-5, 6)), g["splice11_small"])
    assert np.array_equal(O.splicing(x[:64], range(-8, 9)), g["splice17"])
    assert np.array_equal(O.splicing(x[:8], range(0, 1)), g["splice1"])


def test_transform_bit_exact(golden_dir):
    g = _g(golden_dir, "splice.npz")
    ft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    got = O.apply_kaldi_feature_transform(O.splicing(g["x"], range(-5, 6)), ft)
    assert got.dtype == np.float32
    assert np.array_equal(got, g["splice11_ft"])


def test_middle_block_and_time_delay(golden_dir):
    g = _g(golden_dir, "timedelay.npz")
    ft = O.load_kaldi_feature_transform(os.path.join(golden_dir, "final.feature_transform"))
    ftm = O.select_transform_for_network(ft, "lstm")
    assert ftm["shape"][0] == 40 and ftm["shifts"] == [0]
    assert np.array_equal(O.apply_kaldi_feature_transform(g["x"], ftm), g["x_mid_ft"])
    xd, offd = O.apply_time_delay_x(g["x"], g["offsets"], 5)
    assert np.array_equal(xd, g["x_delayed"])
    assert np.array_equal(offd, g["offsets_delayed"])
    fft = O.select_transform_for_network(ft, "tdnn", splice=8)
    assert fft["addShift"].shape == (17 * 40,)
    assert np.array_equal(fft["addShift"][:40], ftm["addShift"])


def test_logsum_and_head(golden_dir):
    g = _g(golden_dir, "head.npz")
    ap = np.load(os.path.join(golden_dir, "log_ap_Kaldi1909.npy"))
    assert np.array_equal(O.logsum(g["y"], axis=1), g["logsum"])
    assert np.array_equal(O.log_softmax(g["y"]), g["logsoftmax"])
    assert np.array_equal(O.head(g["y"], ap), g["head_ap"])


def test_prior_matches_counts(golden_dir):
    ap = np.load(os.path.join(golden_dir, "log_ap_Kaldi1909.npy"))
    txt = open(os.path.join(golden_dir, "ali_train_pdf.counts")).read().replace("[", " ").replace("]", " ")
    c = np.asarray([float(t) for t in txt.split()], dtype=np.float64)
    assert ap.shape == (1, 1909) and ap.dtype == np.float32 and c.shape == (1909,)
    want = np.log((c - 0.5) / np.sum(c - 0.5))
    assert np.abs(ap[0] - want).max() < 1e-5
    assert abs(np.exp(ap.astype(np.float64)).sum() - 1.0) < 1e-5
    # 1,124,823 train frames
    assert int(round(c.sum() - 1909 * 0.5)) == 1124823 or int(c.sum()) == 1124823


def test_lab_format(golden_dir, tmp_path):
    g = _g(golden_dir, "head.npz")
    want = open(os.path.join(golden_dir, "sample.lab"), "rb").read()
    p = tmp_path / "a.lab"
    O.save_bin(str(p), g["logsoftmax"][:4])
    assert p.read_bytes() == want
    rows, cols = np.frombuffer(want[:8], dtype=np.int32)  # the C++ reader uses int32
    assert (rows, cols) == (4, 1909)
    assert np.array_equal(O.load_bin(str(p)), g["logsoftmax"][:4])
