#!/usr/bin/env python3
"""Headline benchmark: acoustic-model frames/sec of the network-output hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4]

A "step" is one pass of the hot path (splice/transform -> network -> log-softmax head) over one
synthetic data set.  Default workload = BASELINE.json configs[1]: 6x2048 ReLU MLP on 440 spliced fMLLR +
100-dim i-vectors -> 1909 pdfs, bf16 tcgen05 GEMMs, TIMIT-train-shaped set (3696 utts, 1,124,823 frames).
With N GPUs every rank processes its own set of that size (weak scaling, no collective on the data path).

  value : frames/s with the inputs already resident in HBM and the outputs left in HBM
  e2e   : frames/s through the public predict() with pinned HOST buffers: H2D of features/i-vectors and
          D2H of the (N, 1909) log-likelihoods inside the timed region
  --impl reference : the reference's CPU forward (NumPy restatement = what Chainer's CPU backend executes;
          Chainer itself is not installable offline), all host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLASSES = 1909
TRAIN_UTTS, TRAIN_FRAMES = 3696, 1124823
TEST_UTTS = 1344

WORKLOADS = {
    # name: (network, in_feat, ivec, units, layers, set, precision, flop/frame)
    "cfg1": dict(network="ff", ivec=0, units=1024, layers=6, utts=TEST_UTTS, frames=None, flop=15296512,
                 desc="cfg1: 6x1024 ReLU MLP, 440 spliced fMLLR -> 1909, test-shaped 1344 utts"),
    "cfg2": dict(network="ff", ivec=100, units=2048, layers=6, utts=TRAIN_UTTS, frames=TRAIN_FRAMES, flop=51974144,
                 desc="cfg2: 6x2048 ReLU MLP, 440 spliced fMLLR + 100 i-vector -> 1909, train-shaped 3696 utts / 1,124,823 frames"),
    "cfg3": dict(network="lstm", ivec=0, units=512, layers=4, utts=TEST_UTTS, frames=None, flop=16798720,
                 desc="cfg3: 4x512 LSTM, 40 fMLLR -> 1909, timedelay 5, test-shaped 1344 utts"),
    "cfg3t": dict(network="lstm", ivec=0, units=512, layers=4, utts=TRAIN_UTTS, frames=TRAIN_FRAMES, flop=16798720,
                  desc="cfg3t: 4x512 LSTM, 40 fMLLR -> 1909, timedelay 5, train-shaped 3696 utts / 1,124,823 frames"),
    "cfg4": dict(network="blstm", ivec=100, units=512, layers=4, utts=TEST_UTTS, frames=None, flop=46999552,
                 desc="cfg4: 4x(2x512) bidirectional LSTM, 40 fMLLR + 100 i-vector -> 1909, test-shaped 1344 utts"),
    "cfg4g": dict(network="bgru", ivec=100, units=512, layers=4, utts=TEST_UTTS, frames=None,
                  flop=2 * (2 * (140 * 1536 + 512 * 1536) + 3 * 2 * (1024 * 1536 + 512 * 1536) + 1024 * 1909),
                  desc="cfg4g: 4x(2x512) bidirectional GRU, 40 fMLLR + 100 i-vector -> 1909, test-shaped 1344 utts"),
    # cfg5: ensembles of 10 fold models, logit mean fused into the head (evaluate.py:35-51)
    "cfg5": dict(network="ff", ivec=0, units=1024, layers=6, utts=TEST_UTTS, frames=None, flop=10 * 15296512, folds=10,
                 desc="cfg5: ensemble of 10 fold 6x1024 MLPs (logit mean), 440 spliced fMLLR -> 1909, test-shaped 1344 utts"),
    "cfg5b": dict(network="blstm", ivec=100, units=512, layers=4, utts=TEST_UTTS, frames=None, flop=10 * 46999552,
                  folds=10,
                  desc="cfg5b: ensemble of 10 fold 4x(2x512) BLSTMs (logit mean), 40 fMLLR + 100 i-vector -> 1909, test-shaped 1344 utts"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def transform():
    from oracle import nnam_oracle as O
    return O.load_kaldi_feature_transform(os.path.join(ROOT, "tests", "golden", "final.feature_transform"))


def make_workload(name, rank):
    from oracle import nnam_oracle as O
    w = WORKLOADS[name]
    x, offsets, iv = O.synth_set(1234 + 17 * rank, w["utts"], 40, w["ivec"], total=w["frames"])
    return w, x, offsets, iv


def make_params(w, seed=4321):
    from oracle import nnam_oracle as O
    rng = np.random.default_rng(seed)
    if w["network"] == "ff":
        return O.init_mlp(rng, 440 + w["ivec"], w["units"], w["layers"], N_CLASSES)
    bid = w["network"] in ("blstm", "bgru")
    cell = "gru" if w["network"] == "bgru" else "lstm"
    return O.init_recurrent(rng, cell, 40 + w["ivec"], w["units"], w["layers"], N_CLASSES, bidirectional=bid)


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML in a background thread DURING the timed
    region (same data as the nvidia-smi clocks line of the profiling recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def _handle(self, nv):
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:  # noqa: BLE001
            return nv.nvmlDeviceGetHandleByIndex(self.index)

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = self._handle(nv)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, mx, pw, rs))
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        import threading
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: " + str(self.err)]}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = set()
        for r in self.rows:
            for b, n in bits.items():
                if r[3] & b:
                    reasons.add(n)
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(self.rows[0][1]),
                "reasons": sorted(reasons), "power_w_max": float(max(r[2] for r in self.rows)), "samples": len(sm)}


def cpu_reference_rate(wname, sample_frames, steps=1, warmup=0):
    """Frames/s of the reference-equivalent NumPy CPU forward (oracle port) on a bounded sample."""
    from oracle import nnam_oracle as O
    w, x, offsets, iv = make_workload(wname, 0)
    folds = w.get("folds", 1)
    ps = [make_params(w, 4321 + k) for k in range(folds)]
    p = ps[0]
    ft = transform()
    if folds > 1:
        sample_frames = max(sample_frames // folds, 2048)
    if w["network"] == "ff":
        n = min(sample_frames, len(x))
        xs, ivs = x[:n], (iv[:n] if iv is not None else None)

        def run():
            # predict_folds.predict FF loop (batch 1024) with the train/evaluate feature order
            ys = []
            for o in range(0, n, 1024):
                e = min(o + 1024, n)
                f = O.apply_kaldi_feature_transform(O.prepare_batch(xs, np.arange(o, e), 11), ft)
                if ivs is not None:
                    f = np.concatenate((f, ivs[o:e]), axis=1)
                if folds > 1:  # evaluate.py:35-51: mean of the fold logits, then log-softmax
                    ys.append(O.log_softmax(O.nn_with_rpl(None, [(lambda v, q=q: O.mlp_forward(q, v, w["layers"]))
                                                                 for q in ps], None, f)))
                else:
                    ys.append(O.log_softmax(O.mlp_forward(p, f, w["layers"])))
            return np.concatenate(ys)
        sample = f"first {n} frames, predict() FF loop batch 1024"
    else:
        n_utt = int(min(max(sample_frames // 1024, 16), 256, len(offsets) - 1))  # the loop batches all of them per step
        bid = w["network"] in ("blstm", "bgru")
        cell = "gru" if w["network"] == "bgru" else "lstm"
        if bid:  # no batched reference loop exists for the bidirectional nets: per-utterance, so a smaller sample
            n_utt = max(n_utt // 4, 16) if folds == 1 else max(n_utt // (4 * folds), 2)
        off = offsets[:n_utt + 1]
        n = int(off[-1])
        ftm = O.select_transform_for_network(ft, "lstm")
        xs = x[:n] if iv is None else np.concatenate((O.apply_kaldi_feature_transform(x[:n], ftm), iv[:n]), axis=1)

        def run():
            if bid:
                return [O.log_softmax(sum(O.birnn_forward_utterance(q, cell, w["layers"], xs[off[i]:off[i + 1]])
                                          for q in ps) / np.float32(folds)) for i in range(n_utt)]
            net = O.RecurrentNet(p, "lstm", w["layers"])
            return O.predict(net, xs, off, "lstm", 1, 5, ftm if iv is None else None)
        sample = f"first {n_utt} utterances ({n} frames), time-major loop"
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, sample, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    rate, dt, sample, n = cpu_reference_rate(args.workload, args.cpu_sample, steps=max(args.steps, 1),
                                             warmup=min(args.warmup, 1))
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "acoustic-model frames/sec", "value": rate, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "NumPy/OpenBLAS restatement of the Chainer CPU forward (Chainer 3.5 not installable offline)"},
        "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import nnacousticmodeling_b200 as nn
    from nnacousticmodeling_b200 import engine, ops

    from nnacousticmodeling_b200 import dist_util
    world, rank, local = dist_util.env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init("nccl", dev)
    numa_cpus = dist_util.bind_to_gpu_numa(local) if world > 1 else None  # before the pinned buffers are allocated
    peaks = load_peaks()

    w, x, offsets, iv = make_workload(args.workload, rank)
    n = len(x)
    ft_full = transform()
    net = w["network"]
    recurrent = nn.is_nn_recurrent(net)
    folds = w.get("folds", 1)
    members = []
    for k in range(folds):
        m = nn.get_nn(net, w["layers"], [w["units"]], N_CLASSES, nn.F.relu, [5])
        m.load_params(make_params(w, 4321 + k))
        m.precision = args.precision
        m.to_gpu(local)
        members.append(m)
    model = members[0] if folds == 1 else members
    head = None if folds == 1 else nn.HeadSpec(weights=[1.0 / folds] * folds)
    ft = nn.adapt_transform(ft_full, net, 5, recurrent)
    timedelay = 5 if net == "lstm" else 0

    # pinned host buffers for the e2e leg; device-resident copies for the kernel leg
    xp = nn.empty_pinned(x.shape)
    xp[:] = x
    ivp = None
    if iv is not None:
        ivp = nn.empty_pinned(iv.shape)
        ivp[:] = iv
    out_host = nn.empty_pinned((n, N_CLASSES))
    x_dev = torch.from_numpy(xp).to(dev)
    iv_dev = None if ivp is None else torch.from_numpy(ivp).to(dev)
    out_dev = torch.empty((n, N_CLASSES), dtype=torch.float32, device=dev)

    if recurrent:
        from nnacousticmodeling_b200 import recurrent_engine

        def step_device():
            recurrent_engine.forward_utterances(model, x_dev, offsets, out_dev, 0, len(offsets) - 1, ft=ft,
                                                ivectors=iv_dev, timedelay=timedelay, device=local, head=head)
    else:
        def step_device():
            engine.ff_forward_frames(model, x_dev, ft, 5, out_dev, ivectors=iv_dev, device=local, head=head)

    def step_e2e():
        nn.predict(model, xp, offsets if recurrent else None, N_CLASSES, net, local, 11, timedelay, ft,
                   progress=False, ivectors=ivp, out=out_host, head=head)

    def barrier():
        torch.cuda.synchronize()
        dist_util.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        return dist_util.max_over_ranks(v, dev)

    # ---- device-resident leg ("value"): inputs in HBM, outputs left in HBM, CUDA events
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sum(ops.LAUNCHES.values())
    ops.PROFILE = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    prof, ops.PROFILE = ops.PROFILE, None
    launches = sum(ops.LAUNCHES.values()) - launches0
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    value = world * n / (dev_ms * 1e-3)

    kern = {}
    for name, s, e, work in prof:
        k = kern.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
        k["launches"] += 1
        k["ms"] += s.elapsed_time(e)
        k["work"] += work
    for name, k in kern.items():
        k["ms_per_step"] = k["ms"] / args.steps
        if name in ("gemm", "rnn"):
            k["tflops"] = k["work"] / (k["ms"] * 1e-3) / 1e12
        else:
            k["gbs"] = k["work"] / (k["ms"] * 1e-3) / 1e9
        del k["work"], k["ms"]
    dom = max(kern, key=lambda a: kern[a]["ms_per_step"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{args.workload}:{dom}")
    if dom in ("gemm", "rnn"):
        roof = {"bound": "tensor", "kernel": "gemm_bias_act_kernel + gemm_bias_act_2sm_kernel" if dom == "gemm" else "rnn_seq_kernel",
                "achieved": kern[dom]["tflops"], "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": kern[dom]["tflops"] / peaks["tf_sustained"], "traffic": traffic,
                "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step); burst peak "
                               f"{peaks['tf_burst']} -> frac {kern[dom]['tflops'] / peaks['tf_burst']:.3f}",
                "launches_per_step": kern[dom]["launches"] // args.steps}
        if dom == "rnn":
            roof["note"] = ("K3 is sequential in t and bound by the per-step exchange latency, not by the tensor pipe "
                            "(SURVEY 8d: no roofline claim); achieved = algorithmic lateral FLOP / kernel time")
    else:
        roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peaks["hbm"], "traffic": traffic, "peak_source": peaks["src"]}

    # ---- end-to-end leg: public predict() with pinned host buffers, H2D + D2H inside the timed region
    h2d = x.nbytes + (0 if iv is None else iv.nbytes)
    d2h = n * N_CLASSES * 4
    if args.no_e2e:  # profiling runs only (ncu): the printed line is then not a bench value
        e2e_s, e2e_value, same = float("nan"), None, None
    else:
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
        barrier()
        e2e_value = world * n / e2e_s
        same = bool(torch.equal(torch.from_numpy(out_host[:4096]).to(dev), out_dev[:4096]))

    line = {
        "metric": "acoustic-model frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "bf16x3(fp32-accurate)"}.get(args.precision, args.precision),
        "data": "synthetic",
        "config": {"workload": w["desc"], "frames_per_gpu": n, "precision": args.precision,
                   "l2": "inputs+activations larger than L2 (no flush needed)", "parallelism": f"dp{world} utterance shards, no collective",
                   "cpu_binding": None if numa_cpus is None else f"{len(numa_cpus)} CPUs local to the GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "matches_device_leg": same},
        "gpu_launches": launches,
        "roofline": roof,
        "kernels": kern,
        "flop_per_frame": w["flop"],
        "model_tflops": value * w["flop"] / 1e12,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, dt, sample, _ = cpu_reference_rate(args.workload, args.cpu_sample)
        line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": sample, "seconds": dt}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    dist_util.finalize()
    return 0


_RESULT_OUT = None


def _claim_stdout():
    """Rank 0 prints ONE JSON line on stdout.  Libraries also write there (NCCL prints its version banner from C code
    at communicator creation, whatever NCCL_DEBUG_FILE says), so the real stdout is kept aside for the result line and
    file descriptor 1 points at stderr for everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16",
                    help="bf16 | fp16 | fp32 (bf16x3) | bf16+a[:layers][+w[:layers]] -- see engine.Precision")
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="frames of the workload timed on the CPU (default: ~10-30 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu only)")
    args = ap.parse_args()
    if args.cpu_sample is None:
        # the cpu_baseline leg runs once (~10-30 s); the reference arm repeats its sample steps + warmup times
        args.cpu_sample = 65536 if args.impl == "reference" else 262144
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
