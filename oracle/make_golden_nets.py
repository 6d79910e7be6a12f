#!/usr/bin/env python3
"""Generate tests/golden/nets/*.npz by running the reference's OWN hot-path code, unmodified.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference); the GPU box uses the committed
fixtures.  The reference's scripts need Chainer 3.5, which cannot be installed offline; ``tests/chainer_shim`` supplies
the few Chainer classes / functions they touch as NumPy code (what Chainer's CPU backend computes), so these files run
as they are:

  scripts/common/chainer_networks.py   model topologies, parameter names (get_nn: all nine kinds)
  scripts/common/MGRU.py               the MGRU / GRU step incl. the first-step rule (:67-85)
  scripts/common/RPL.py                RPL4 (:58-74)
  scripts/common/predict_folds.py      predict() and main(): FF chunking, RNN time-major loop, time delay, fold / dev mode
  scripts/common/evaluate.py           NNWithRPL (:19-51), main(): splice -> transform -> i-vector order
  scripts/util/evaluateModelForTest.py evaluateModelTestTri (:36-122), .lab files
  scripts/common/master_script.py      main(): the command lists handed to predict_folds.main / evaluate.main

Two families of fixtures (format: oracle/golden_layout.py):
  cells_<kind>.npz    predict_folds.main on a 2-fold + dev tree, 39 classes, one per get_nn kind
  replay_<name>.npz   master_script.main (--no-train-*) on a full tree, 1909 classes: its predict_folds.main calls
                      (fold + dev mode) and all its evaluate.main calls (folds / master / rpl combinations)

The only accommodations, none of them an edit of a reference file: ``np.int = int`` (predict_folds.py:31 and
evaluateModelForTest.py:55 use the alias NumPy removed in 1.24), and ``master_script.evaluate_main`` is wrapped to
record its argument list, snapshot the .lab files each call writes, and swallow the FileNotFoundError of the absent
PhoneRecog binary (evaluateModelForTest.py:124-127, out of scope).

  python oracle/make_golden_nets.py [--out tests/golden/nets]
"""
import argparse
import io
import json
import os
import shutil
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/scripts"
GOLDEN = os.path.join(ROOT, "tests", "golden")

np.int = int  # noqa: NPY001  (see module docstring)
sys.path[:0] = [os.path.join(ROOT, "tests", "chainer_shim"), os.path.join(REF, "common"), os.path.join(REF, "util"), ROOT]

import chainer  # noqa: E402  (the shim)
import chainer.functions as F  # noqa: E402
import chainer.links as L  # noqa: E402
import chainer_networks  # noqa: E402  (reference)
import evaluate as ref_evaluate  # noqa: E402  (reference)
import master_script  # noqa: E402  (reference)
import predict_folds  # noqa: E402  (reference)
from RPL import RPL4  # noqa: E402  (reference)

from oracle.golden_layout import load_lab, pack_text  # noqa: E402

F32 = np.float32


# --------------------------------------------------------------------------------------------------------------------
def seeded_params(link, rng, in_dim):
    """Give every parameter of a reference-built link deterministic values: LeCun-normal weights, N(0, 0.1) biases (so
    that every bias path -- GRU's hidden-side biases, the LSTM forget bias -- is observable)."""
    chainer.config.train = False
    model = link.predictor
    if not isinstance(model, RPL4):
        # the links are built with in_size=None: calling the model fixes the shapes (twice: the GRU family's W_r is
        # only reached once there is a hidden state, MGRU.py:70-74)
        model(np.zeros((1, in_dim), dtype=F32))
        model(np.zeros((1, in_dim), dtype=F32))
        if hasattr(model, "reset_state"):
            model.reset_state()
    for path, prm in link.namedparams():
        shape = prm.data.shape
        leaf = path.rsplit("/", 1)[1]
        if isinstance(model, RPL4):
            val = {"W": 0.05 * rng.standard_normal(shape), "b": 0.1 * rng.standard_normal(shape),
                   "lb": -8.0 + rng.standard_normal(shape)}[leaf]
        elif leaf == "W":
            val = rng.standard_normal(shape) / np.sqrt(np.prod(shape[1:]))
        else:
            val = 0.1 * rng.standard_normal(shape)
        # values are rounded to fp16 so that the fixture can store them in half the bytes (materialise() widens
        # them back to the float32 arrays Chainer's npz files hold)
        prm.data = val.astype(np.float16).astype(F32)


def net_arrays(spec, num_classes, rng, in_dim):
    model = chainer_networks.get_nn(spec["network"], spec["layers"], spec["units"], num_classes, F.relu,
                                    spec.get("tdnn_ksize", [5]), spec.get("dropout", [0]))
    cls = L.Classifier(model)
    seeded_params(cls, rng, in_dim)
    return {path.lstrip("/"): prm.data for path, prm in cls.namedparams()}


def spec_args(spec):
    a = ["-n", spec["network"], "-l", spec["layers"], "-u"] + list(spec["units"])
    if spec.get("splice"):
        a += ["--splice", spec["splice"]]
    if spec.get("timedelay"):
        a += ["--timedelay", spec["timedelay"]]
    if "tdnn_ksize" in spec:
        a += ["--tdnn-ksize"] + list(spec["tdnn_ksize"])
    # predict_folds.py:108 defaults --dropout to the int 0, which get_nn then indexes (chainer_networks.py:165): the
    # reference only runs when -d is given, as in master_script.py:29's default network spec
    a += ["-d"] + list(spec.get("dropout", [0]))
    return [str(t) for t in a]


def synth(rng, lens, dim=40):
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return rng.standard_normal((int(off[-1]), dim)).astype(F32), off


class Tree:
    """Collects the fixture entries while writing the same files under ``root`` for the reference to read."""

    def __init__(self, root):
        self.root, self.entries = root, {}

    def _path(self, rel):
        p = os.path.join(self.root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        return p

    def npy(self, rel, arr):
        np.save(self._path(rel), arr)
        self.entries["in:" + rel] = arr

    def archive(self, rel, arrays):
        with open(self._path(rel), "wb") as f:
            np.savez(f, **arrays)
        for k, v in arrays.items():
            assert np.array_equal(v.astype(np.float16).astype(F32), v)
            self.entries[f"npz:{rel}::{k}"] = v.astype(np.float16)

    def text(self, rel, s):
        with open(self._path(rel), "w") as f:
            f.write(s)
        self.entries["txt:" + rel] = pack_text(s)

    def copy(self, rel, golden_name):
        shutil.copy(os.path.join(GOLDEN, golden_name), self._path(rel))
        self.entries["copy:" + rel] = pack_text(golden_name)

    def save(self, out_file, meta, outputs):
        e = dict(self.entries)
        e["meta"] = pack_text(json.dumps(meta))
        for k, v in outputs.items():
            e["out:" + k] = v
        np.savez_compressed(out_file, **e)


def in_dim_of(spec, ivec_dim=0):
    recurrent = chainer_networks.is_nn_recurrent(spec["network"])
    if spec["network"] == "tdnn":
        win = sum(spec["tdnn_ksize"]) - len(spec["tdnn_ksize"]) + 1
        return 40 * win + ivec_dim
    return (40 if recurrent else 40 * (2 * spec.get("splice", 0) + 1)) + ivec_dim


# --------------------------------------------------------------------------------------------------------------------
CELL_SPECS = {
    "ff": dict(network="ff", layers=2, units=[48], splice=5),
    "tdnn": dict(network="tdnn", layers=2, units=[24, 16], tdnn_ksize=[3, 3]),
    "lstm": dict(network="lstm", layers=2, units=[64], timedelay=3),
    "zoneoutlstm": dict(network="zoneoutlstm", layers=1, units=[64], timedelay=2, dropout=[0.5, 0.5]),
    "zoneoutdropoutlstm": dict(network="zoneoutdropoutlstm", layers=1, units=[64], timedelay=2, dropout=[0.2, 0.5, 0.5]),
    "peepholelstm": dict(network="peepholelstm", layers=2, units=[64], timedelay=3),
    "gru": dict(network="gru", layers=2, units=[64], timedelay=3),
    "mgrurelu": dict(network="mgrurelu", layers=1, units=[64], timedelay=1),
    "mgrurelur": dict(network="mgrurelur", layers=2, units=[64]),
}


def make_cells(kind, spec, out_dir, seed):
    """predict_folds.main in fold mode and dev mode (the two calls of master_script.py:177-211, without --tri)."""
    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as root:
        t = Tree(root)
        t.copy("data/fmllr/final.feature_transform", "final.feature_transform")
        fold_lens = [[23, 9, 31], [14, 27, 8, 11]]
        if kind not in ("ff", "tdnn", "lstm", "mgrurelur"):
            fold_lens = fold_lens[:1]  # one fold keeps the fixture small; the fold loop is pinned by the other kinds
        for k, lens in enumerate(fold_lens):
            x, off = synth(rng, lens)
            t.npy(f"folds_in/data_{k}.npy", x)
            t.npy(f"folds_in/offsets_{k}.npy", off)
            t.archive(f"models/fold_{k}.npz", net_arrays(spec, 39, rng, in_dim_of(spec)))
        x, off = synth(rng, [17, 12])
        t.npy("data/fmllr/data_dev.npy", x)
        t.npy("data/offsets_dev.npy", off)
        base = ["--ft", "final.feature_transform", "--gpu", "-1", "--no-progress", "--fold-model-dir", "models",
                "--fold-output-dir", "folds_out"] + spec_args(spec)
        cmds = {"fold": base + ["--fold-data-dir", "folds_in"],
                "dev": base + ["--data-dir", "data/fmllr", "--offset-dir", "data", "--fold-output-dev", "data_dev.npy"]}
        # fold mode reads the transform from --data-dir too (predict_folds.py:171)
        cmds["fold"] += ["--data-dir", "data/fmllr"]
        cwd = os.getcwd()
        os.chdir(root)
        try:
            with redirect_stdout(io.StringIO()):
                predict_folds.main(cmds["fold"])
                predict_folds.main(cmds["dev"])
        finally:
            os.chdir(cwd)
        outs = {f"folds_out/data_{k}.npy": np.load(os.path.join(root, f"folds_out/data_{k}.npy"))
                for k in range(len(fold_lens))}
        outs["folds_out/data_dev.npy"] = np.load(os.path.join(root, "folds_out/data_dev.npy"))
        # the product rejects --gpu -1 (no CPU path): the replayed lists carry the device the test runs on instead
        for c in cmds.values():
            i = c.index("--gpu")
            c[i + 1] = "{GPU}"
        meta = dict(kind=kind, spec=spec, num_classes=39, n_folds=len(fold_lens), cmds=cmds)
        t.save(os.path.join(out_dir, f"cells_{kind}.npz"), meta, outs)
    return {k: v.shape for k, v in outs.items()}


# --------------------------------------------------------------------------------------------------------------------
REPLAYS = {
    # name: (network spec, i-vector dim, folds, extra master_script flags)
    "ff": (dict(network="ff", layers=2, units=[8], splice=5), 0, 2, []),
    "lstm": (dict(network="lstm", layers=1, units=[64], timedelay=2), 0, 1, []),
    # with i-vectors the reference's predict_folds step cannot run with --ft, so these two replay
    # the evaluation stage only
    "mgrurelur_ivec": (dict(network="mgrurelur", layers=1, units=[64]), 10, 1, ["--no-predict"]),
    "ff_ivec_master": (dict(network="ff", layers=2, units=[8], splice=5), 10, 2, ["--no-predict", "--eval-only-master"]),
}


def make_replay(name, spec, ivec_dim, n_folds, extra, out_dir, seed):
    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as root:
        t = Tree(root)
        t.copy("data/fmllr/final.feature_transform", "final.feature_transform")
        t.copy("recog/log_ap_Kaldi1909.npy", "log_ap_Kaldi1909.npy")
        fold_dir = f"results/fold_data/{n_folds}/fmllr" + ("+ivec_train" if ivec_dim else "")
        for k, lens in enumerate([[4, 3], [3, 5]][:n_folds]):
            x, off = synth(rng, lens)
            t.npy(f"{fold_dir}/data_{k}.npy", x)
            t.npy(f"{fold_dir}/offsets_{k}.npy", off)
            if ivec_dim:
                t.npy(f"{fold_dir}/ivectors_{k}.npy", (0.5 * rng.standard_normal((len(x), ivec_dim))).astype(F32))
            t.archive(f"results/models/folds/{n_folds}/tmp/fold_{k}.npz", net_arrays(spec, 1909, rng, in_dim_of(spec, ivec_dim)))
        t.archive(f"results/models/master/{n_folds}/tmp/model", net_arrays(spec, 1909, rng, in_dim_of(spec, ivec_dim)))
        rpl_cls = L.Classifier(RPL4(1909))
        seeded_params(rpl_cls, rng, 0)
        t.archive(f"results/models/rpl/{n_folds}/tmp/model", {p.lstrip("/"): q.data for p, q in rpl_cls.namedparams()})
        for part, lens in (("dev", [4, 3]), ("test", [4, 3])):
            x, off = synth(rng, lens)
            t.npy(f"data/fmllr/data_{part}.npy", x)
            t.npy(f"data/offsets_{part}.npy", off)
            if ivec_dim:
                d = "ivec_train" if part == "dev" else "ivec_test"
                t.npy(f"data/{d}/ivectors_{part}.npy", (0.5 * rng.standard_normal((len(x), ivec_dim))).astype(F32))
        t.text("data/test.list", "fadg0_si1279\nmbpm0_sx317\n")
        argv = ["master_script.py", "--num-folds", str(n_folds), "--data-dir", "data/fmllr", "--offset-dir", "data",
                "--utt-list-dir", "data", "--recog-dir", "recog", "--output-dir", "results",
                "--network-spec", " ".join(spec_args(spec)), "--no-train-master", "--no-train-folds", "--no-train-rpl",
                "--no-progress", "--eval-data", "test"] + extra
        if ivec_dim:
            argv += ["--ivector-dir", "data/ivec_train", "data/ivec_test"]

        calls, outs = [], {}
        real_predict, real_eval = master_script.predict_folds_main, master_script.evaluate_main

        def predict_wrapper(cmd):
            calls.append(["predict_folds"] + [str(c) for c in cmd])
            real_predict(cmd)

        def evaluate_wrapper(cmd):
            tag = f"eval{sum(1 for c in calls if c[0] == 'evaluate')}"
            calls.append(["evaluate"] + [str(c) for c in cmd])
            shutil.rmtree("lab", ignore_errors=True)
            try:
                real_eval(cmd)
            except FileNotFoundError as e:  # recog/PhoneRecog (evaluateModelForTest.py:124-127) is out of scope
                if "PhoneRecog" not in str(e.filename):
                    raise
            for line in open("data/test.list"):
                outs[f"{tag}/{line.strip()}.lab"] = load_lab(os.path.join("lab", line.strip() + ".lab"))

        cwd, old_argv = os.getcwd(), sys.argv
        os.chdir(root)
        master_script.predict_folds_main, master_script.evaluate_main = predict_wrapper, evaluate_wrapper
        sys.argv = argv
        try:
            with redirect_stdout(io.StringIO()):
                master_script.main()
        finally:
            sys.argv = old_argv
            master_script.predict_folds_main, master_script.evaluate_main = real_predict, real_eval
            os.chdir(cwd)
        out_dir_rel = f"results/fold_data_out/{n_folds}/tmp"
        if "--no-predict" not in extra:
            for k in range(n_folds):
                outs[f"{out_dir_rel}/data_{k}.npy"] = np.load(os.path.join(root, out_dir_rel, f"data_{k}.npy"))
            outs[f"{out_dir_rel}/data_dev.npy"] = np.load(os.path.join(root, out_dir_rel, "data_dev.npy"))
        meta = dict(name=name, spec=spec, num_classes=1909, ivec_dim=ivec_dim, n_folds=n_folds, master_script_argv=argv[1:], calls=calls)
        t.save(os.path.join(out_dir, f"replay_{name}.npz"), meta, outs)
    return calls, {k: v.shape for k, v in outs.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(GOLDEN, "nets"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for i, (kind, spec) in enumerate(CELL_SPECS.items()):
        print("cells", kind, make_cells(kind, spec, args.out, 100 + i))
    for i, (name, (spec, ivec, n_folds, extra)) in enumerate(REPLAYS.items()):
        calls, shapes = make_replay(name, spec, ivec, n_folds, extra, args.out, 200 + i)
        print("replay", name, [c[0] for c in calls], len(shapes), "outputs")


if __name__ == "__main__":
    main()
