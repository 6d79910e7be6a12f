#!/usr/bin/env python3
"""Parity table of the precision modes at BASELINE.json's full sizes (run on the B200 box).

For every (workload, precision mode, weight variant) it runs the product path over the WHOLE synthetic set and compares a
large random sample of frames / utterances with the fp32 oracle:

  max_abs     max |log-likelihood - oracle|                      (north_star: <= 1e-3 fp32 mode, <= 5e-2 16-bit modes)
  agree_raw   fraction of frames whose argmax equals the oracle's (north_star: >= 0.995) -- no near-tie allowance
  agree_near  (diagnostic only) the class we pick is within 1e-2 of the oracle's best in the oracle's own scores

Weight variants: "plain" = fp32 LeCun-normal weights as drawn; "e16" = the same weights pre-rounded to the mode's 16-bit
element type ON BOTH SIDES (SURVEY 8d "bf16-exact weights" variant: identical weights, the oracle stays fp32 arithmetic).

  python tests/tools/gpu_parity_table.py --workloads cfg2,cfg3 --modes bf16,fp16 --out gpurun_out/parity.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import nnam_oracle as O  # noqa: E402  (the checker)
import bench  # noqa: E402  (workload table + generators)


def round_e16(w, elem):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
    dt = torch.float16 if elem == "fp16" else torch.bfloat16
    return t.to(dt).to(torch.float32).numpy()


def elem_of(mode):
    return "fp16" if mode.startswith("fp16") else "bf16"


def stats(got, want):
    rows = np.arange(len(want))
    pick = got.argmax(axis=1)
    return dict(max_abs=float(np.abs(got - want).max()),
                agree_raw=float(np.mean(pick == want.argmax(axis=1))),
                agree_near=float(np.mean(want[rows, pick] >= want.max(axis=1) - 1e-2)),
                frames=int(len(want)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="cfg2,cfg3,cfg4")
    ap.add_argument("--modes", default="bf16")
    ap.add_argument("--weights", default="plain,e16")
    ap.add_argument("--ff-sample", type=int, default=20000)
    ap.add_argument("--rnn-sample", type=int, default=40)
    ap.add_argument("--birnn-sample", type=int, default=12)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.jsonl"))
    args = ap.parse_args()

    import nnacousticmodeling_b200 as nn
    ft_full = nn.loadKaldiFeatureTransform(os.path.join(ROOT, "tests", "golden", "final.feature_transform"))
    oft_full = O.load_kaldi_feature_transform(os.path.join(ROOT, "tests", "golden", "final.feature_transform"))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a")

    for wname in args.workloads.split(","):
        w, x, offsets, iv = bench.make_workload(wname, 0)
        net = w["network"]
        recurrent = nn.is_nn_recurrent(net)
        bid = net in ("blstm", "bgru")
        cell = "gru" if net == "bgru" else "lstm"
        timedelay = 5 if net == "lstm" else 0
        p_plain = bench.make_params(w)
        rng = np.random.default_rng(0)
        oracle_cache = {}
        if recurrent:
            lens = np.diff(offsets)
            k = args.birnn_sample if bid else args.rnn_sample
            utts = sorted({int(lens.argmin()), int(lens.argmax()), *rng.integers(0, len(lens), k).tolist()})
        else:
            idx = np.unique(np.concatenate([np.arange(8), np.arange(len(x) - 8, len(x)),
                                            rng.integers(0, len(x), args.ff_sample)]))
        for mode in args.modes.split(","):
            for wv in args.weights.split(","):
                elem = elem_of(mode)
                if wv == "plain":
                    p, okey = p_plain, "plain"
                else:
                    p = {k2: (round_e16(v, elem) if k2.endswith("/W") else v) for k2, v in p_plain.items()}
                    okey = "e16-" + elem
                # ---- oracle (fp32 arithmetic on the same weights), cached per weight set
                if okey not in oracle_cache:
                    t0 = time.time()
                    if not recurrent:
                        feats = O.apply_kaldi_feature_transform(O.prepare_batch(x, idx, 11), oft_full)
                        if iv is not None:
                            feats = np.concatenate((feats, iv[idx]), axis=1)
                        want = O.log_softmax(O.mlp_forward(p, feats, w["layers"]))
                        rows = idx
                    else:
                        oft = O.select_transform_for_network(oft_full, "lstm")
                        wants, rws = [], []
                        if bid:
                            for u in utts:
                                xs = O.apply_kaldi_feature_transform(x[offsets[u]:offsets[u + 1]], oft)
                                if iv is not None:
                                    xs = np.concatenate((xs, iv[offsets[u]:offsets[u + 1]]), axis=1)
                                wants.append(O.log_softmax(O.birnn_forward_utterance(p, cell, w["layers"], xs)))
                                rws.append(np.arange(offsets[u], offsets[u + 1]))
                        else:
                            # the reference's time-major loop over the sampled utterances (timedelay, quirk Q4)
                            sub_off = np.concatenate([[0], np.cumsum([lens[u] for u in utts])])
                            xs = np.concatenate([x[offsets[u]:offsets[u + 1]] for u in utts])
                            y = O.predict(O.RecurrentNet(p, net, w["layers"]), xs, sub_off, net, 1, timedelay, oft)
                            for i, u in enumerate(utts):
                                keep = lens[u] - timedelay  # the tail rows stay 0 on both sides (checked elsewhere)
                                wants.append(y[sub_off[i]:sub_off[i] + keep])
                                rws.append(np.arange(offsets[u], offsets[u] + keep))
                        want, rows = np.concatenate(wants), np.concatenate(rws)
                    oracle_cache[okey] = (want, rows, time.time() - t0)
                want, rows, osec = oracle_cache[okey]
                # ---- product path over the whole set
                m = nn.get_nn(net, w["layers"], [w["units"]], 1909, nn.F.relu, [5])
                m.load_params(p)
                m.precision = mode
                ft = nn.adapt_transform(ft_full, net, 5, recurrent)
                t0 = time.time()
                got = nn.predict(m, x, offsets if recurrent else None, 1909, net, 0, 11, timedelay, ft, progress=False,
                                 ivectors=iv)
                sec = time.time() - t0
                rec = dict(workload=wname, mode=mode, weights=wv, **stats(got[rows], want), predict_s=sec, oracle_s=osec)
                print(json.dumps(rec), flush=True)
                fout.write(json.dumps(rec) + "\n")
                fout.flush()
                del got, m


if __name__ == "__main__":
    main()
