"""Evaluate-side callers of the hot path: ``NNWithRPL`` (scripts/common/evaluate.py:19-51) and the forward +
``.lab`` half of ``evaluateModelTestTri`` (scripts/util/evaluateModelForTest.py:36-122), on the B200 kernels.

The reference evaluates the K+1 networks one after another on the host and averages their raw logits; here
every member runs on the device and K4 fuses the weighted logit mean, RPL4 (RPL.py:68-74), the prior
subtraction ``y - ap`` and the log-softmax (evaluateModelForTest.py:75-77,110-112) into one pass.  The
recurrent branch, like the reference's, applies NO time-delay compensation (quirk Q4: ``--timedelay`` is parsed at
evaluate.py:61 and never used).  Decoding (PhoneRecog) and PER scoring are out of scope: when the reference's
``PhoneRecog`` binary is present in ``recogdir`` it is invoked exactly as the reference does, otherwise the
``.lab`` files and the ``.scp`` list are the result.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

import numpy as np

from . import engine
from ._native import NnamError
from .features import saveBin
from .networks import RPL4


class NNWithRPL:
    """evaluate.py:19-51: master and/or fold networks averaged in the LOGIT domain, optional RPL4 on top."""

    def __init__(self, master=None, folds=(), rpl=None):
        self.master, self.folds, self.rpl = master, list(folds), rpl
        self.num_folds = len(self.folds)
        if master is None and not self.folds:
            raise NnamError("NNWithRPL needs a master network or at least one fold network")
        if rpl is not None and not isinstance(rpl, RPL4):
            raise NnamError("NNWithRPL: only RPL4 is used by the reference's evaluate.py and supported here")

    @property
    def members(self):
        return ([self.master] if self.master is not None else []) + self.folds

    @property
    def weights(self):
        """Weights that reproduce evaluate.py:36-47: master only 1; (K*master + sum folds) / 2K; sum folds / K."""
        k = self.num_folds
        if self.master is not None and k == 0:
            return [1.0]
        if self.master is not None:
            return [0.5] + [1.0 / (2 * k)] * k
        return [1.0 / k] * k

    @property
    def recurrent(self):
        return self.members[0].recurrent

    @property
    def n_out(self):
        return self.members[0].n_out

    def to_gpu(self, device=None):
        for m in self.members:
            m.to_gpu(device)
        return self

    def to_cpu(self):
        raise NnamError("nnacousticmodeling_b200 has no CPU path (GPUID < 0 is not supported)")

    def reset_state(self):
        for m in self.members:
            m.reset_state()

    def head_spec(self, ap=None):
        return engine.HeadSpec(prior=ap, prior_scale=1.0, rpl=None if self.rpl is None else self.rpl.params,
                               weights=self.weights)

    def __call__(self, x):
        """(B, D_in) -> (B, C): averaged logits (after RPL4 when present); one time step for recurrent members."""
        import torch
        from . import ops
        outs = [m(x) for m in self.members]
        if self.rpl is None and len(outs) == 1:
            return outs[0]
        dev = engine._device(self.members[0]._device)
        with torch.cuda.device(dev):
            ts = [o if isinstance(o, torch.Tensor) else torch.from_numpy(o).to(dev) for o in outs]
            rpl = None if self.rpl is None else tuple(engine._dev_vec(self.rpl.params[k], dev) for k in ("W", "b", "lb"))
            y = ops.head([t.contiguous() for t in ts], self.n_out, weights=self.weights, rpl=rpl, final_normalize=False)
            return y if isinstance(outs[0], torch.Tensor) else y.cpu().numpy()


def _as_members(model):
    if isinstance(model, NNWithRPL):
        return model.members, model.head_spec
    return [model], (lambda ap=None: engine.HeadSpec(prior=ap, prior_scale=1.0))


def evaluate_forward(model, data, offsets, ap=None, GPUID=0, rnn=False, out=None):
    """evaluateModelForTest.py:52-122 up to the decoder: (N, C) float32 = log_softmax(model(data) - ap) for every
    frame, rows in the original utterance order (what the reference hands to saveBin utterance by utterance).
    ``data`` is the already spliced / transformed / i-vector-extended matrix of evaluate.py:163-171."""
    if GPUID is None or (not isinstance(GPUID, (list, tuple)) and int(GPUID) < 0):
        raise NnamError("GPUID < 0 (CPU) is not supported by nnacousticmodeling_b200")
    members, spec = _as_members(model)
    data = np.ascontiguousarray(data, dtype=np.float32)
    n_out = members[0].n_out
    ap = None if ap is None else np.asarray(ap, dtype=np.float32).reshape(-1)
    if ap is not None and ap.shape[0] != n_out:
        raise NnamError(f"evaluate: prior has {ap.shape[0]} entries, the network has {n_out} outputs")
    from .predict import predict
    return predict(members if len(members) > 1 else members[0], data, offsets if rnn else None, n_out,
                   members[0].network, GPUID, 1, 0, None, progress=False, out=out, head=spec(ap), presliced=True)


def evaluateModelTestTri(model, data, offsets, PIP, LMW, ap=None, GPUID=0, testOrDev="test", tmpDir="lab",
                         uttlistdir=".", recogdir=".", progress=True, rnn=False):
    """evaluateModelForTest.py:36-134.  Writes one ``.lab`` per utterance (uint32 rows, uint32 cols, float32
    row-major -- kw_utils.py:4-12) and the ``.scp`` list; runs PhoneRecog + scoring only if the reference's binary
    and scoring helpers are available (returns the PER), else returns None after writing the files."""
    with open(os.path.join(uttlistdir, testOrDev + ".list")) as fid:
        test_list = fid.readlines()
    if len(test_list) != len(offsets) - 1:
        print("Error: wrong number of utterances")
        return -1
    lab_dir = Path(tmpDir)
    lab_dir.mkdir(exist_ok=True, parents=True)
    if progress:
        print("Calculating network outputs")
    y = evaluate_forward(model, data, offsets, ap=ap, GPUID=GPUID, rnn=rnn)
    if progress:
        print("Writing output files")
    scp = Path(lab_dir, testOrDev + ".scp")
    with open(str(scp), "wt") as fscp:
        for i, f in enumerate(test_list):
            labout = Path(lab_dir, f.strip() + ".lab")
            labout.parent.mkdir(exist_ok=True, parents=True)
            saveBin(str(labout), y[offsets[i]:offsets[i + 1], :])
            fscp.write(str(labout) + "\n")
    exe = Path(recogdir, "PhoneRecog.exe" if os.name == "nt" else "PhoneRecog")
    if not exe.is_file():
        return None  # decoder is out of scope; the .lab/.scp files are the product of this step
    res = Path(lab_dir, "vysledek_" + testOrDev + ".txt")
    subprocess.run([str(exe), str(scp), str(Path(recogdir, "kaldiTri1909.img")), str(res), str(-abs(PIP)), str(LMW)],
                   cwd=os.getcwd())
    try:  # scoring uses the reference's own helpers when they are importable (not part of this package)
        from evaluateModelForTest import convert_results  # type: ignore
        from kw_utils import loadMlf  # type: ignore
        from levenshtein import computeWER  # type: ignore
    except ImportError:
        return None
    p39 = Path(lab_dir, "vysledek_" + testOrDev + "_p39.txt")
    convert_results(str(Path(recogdir, "phones.60-48-39.map")), str(res), str(p39))
    return computeWER(loadMlf(str(p39)), loadMlf(str(Path(recogdir, testOrDev + "_ref.mlf"))), True)


def main(arg_list=None):
    """scripts/common/evaluate.py:53-214 -- same flags and data flow (splice -> transform -> i-vector concat on the
    device, evaluate.py:163-171), model / NNWithRPL construction (:101-138), then evaluateModelTestTri.  Extensions:
    ``--precision fp32|fp16|bf16`` and ``--tmp-dir`` for the .lab directory (the reference hard-codes 'lab')."""
    import argparse

    from . import functions as F
    from .features import adapt_transform, loadKaldiFeatureTransform, splice_and_transform
    from .networks import Classifier, get_nn, is_nn_recurrent, load_npz

    parser = argparse.ArgumentParser(description="B200 evaluation step (evaluate.py drop-in)")
    parser.add_argument("--network", "-n", default="ff")
    parser.add_argument("--model", "-m", default="")
    parser.add_argument("--units", "-u", type=int, nargs="+", default=[1024])
    parser.add_argument("--layers", "-l", type=int, default=2)
    parser.add_argument("--activation", "-a", default="relu")
    parser.add_argument("--tdnn-ksize", type=int, nargs="+", default=[5])
    parser.add_argument("--timedelay", type=int, default=0)  # parsed and never used, as in the reference (quirk Q4)
    parser.add_argument("--splice", type=int, default=0)
    parser.add_argument("--dropout", "-d", type=float, nargs="+", default=[0])
    parser.add_argument("--tri", action="store_true")
    parser.add_argument("--ft", default="final.feature_transform")
    parser.add_argument("--data-dir", default="data/fmllr")
    parser.add_argument("--offset-dir", default="data")
    parser.add_argument("--ivector-dir")
    parser.add_argument("--recog-dir", required=True)
    parser.add_argument("--utt-list-dir", default="data")
    parser.add_argument("--data", default="data_{}.npy")
    parser.add_argument("--offsets", default="offsets_{}.npy")
    parser.add_argument("--ivectors", default="ivectors_{}.npy")
    parser.add_argument("--PIP", type=float, default=20)
    parser.add_argument("--LMW", type=float, default=1)
    parser.add_argument("--ap-coef", type=float, default=1)
    parser.add_argument("--ap-file", default="log_ap_Kaldi1909.npy")
    parser.add_argument("--gpu", "-g", type=int, default=0)
    parser.add_argument("--test-or-dev", default="test")
    parser.add_argument("--rpl", action="store_true")
    parser.add_argument("--no-rpl-layer", action="store_true")
    parser.add_argument("--rpl-model", default="result_rpl/model")
    parser.add_argument("--fold-model-dir", default="fold_models")
    parser.add_argument("--fold-network-pattern", default="fold_{0}.npz")
    parser.add_argument("--master-network", default="-")
    parser.add_argument("--no-progress", action="store_true")
    parser.add_argument("--precision", default=None, help="fp32 | fp16 | bf16 (engine.Precision)")
    parser.add_argument("--tmp-dir", default="lab")
    args = parser.parse_args(list(map(str, arg_list)) if arg_list is not None else None)

    num_classes = 1909 if args.tri else 39
    if args.activation not in ("sigmoid", "tanh", "relu"):
        print("Wrong activation function specified")
        return None
    activation = F.resolve(args.activation)

    def new_net(path):
        m = get_nn(args.network, args.layers, args.units, num_classes, activation, args.tdnn_ksize, args.dropout)
        if args.precision:
            m.precision = args.precision
        load_npz(str(path), Classifier(m))
        return m

    if args.rpl:
        master = None
        if args.master_network != "-":
            print("Loading master network")
            master = new_net(args.master_network)
        folds = []
        if args.fold_network_pattern != "-":
            while True:
                f = Path(args.fold_model_dir, args.fold_network_pattern.format(len(folds)))
                if not f.is_file():
                    break
                print("Loading fold {} network".format(len(folds)))
                folds.append(new_net(f))
        rpl = None
        # --no-rpl-layer is parsed and never used by the reference (evaluate.py:82,126-131): same here
        if args.rpl_model != "-":
            rpl = RPL4(num_classes)
            with np.load(str(args.rpl_model)) as z:
                rpl.load_params({k: z[k] for k in z.files})
        model = NNWithRPL(master, folds, rpl)
    else:
        model = new_net(args.model)

    recurrent = is_nn_recurrent(args.network)
    splice = (sum(args.tdnn_ksize) - len(args.tdnn_ksize)) // 2 if args.network == "tdnn" else args.splice
    ft = None
    if args.ft is not None and args.ft != "-":
        ft = adapt_transform(loadKaldiFeatureTransform(str(Path(args.data_dir, args.ft))), args.network, splice,
                             recurrent)
    data = np.load(str(Path(args.data_dir, args.data.format(args.test_or_dev))))
    ivectors = None
    if args.ivector_dir is not None:
        ivectors = np.load(str(Path(args.ivector_dir, args.ivectors.format(args.test_or_dev))))
    # evaluate.py:163-171: splicing -> applyKaldiFeatureTransform -> concatenate i-vectors, here one K1 pass
    if splice > 0 or ft is not None or ivectors is not None:
        data = splice_and_transform(data, splice, ft, ivectors, device=args.gpu)
    offsets = np.load(str(Path(args.offset_dir, args.offsets.format(args.test_or_dev))))
    if not args.tri:
        print("Monophones not implemented")
        return None
    ap = np.float32(args.ap_coef) * np.load(str(Path(args.recog_dir, args.ap_file))).astype(np.float32)
    per = evaluateModelTestTri(model, data, offsets, args.PIP, args.LMW, ap=ap, GPUID=args.gpu,
                               testOrDev=args.test_or_dev, tmpDir=args.tmp_dir, uttlistdir=args.utt_list_dir,
                               recogdir=args.recog_dir, progress=not args.no_progress, rnn=recurrent)
    if per is None:
        print("Network outputs written to {}; PhoneRecog is not part of this package".format(args.tmp_dir))
    elif per != -1:
        print("PER: {0:.2f} %".format(per))
    return per


if __name__ == "__main__":
    main()
