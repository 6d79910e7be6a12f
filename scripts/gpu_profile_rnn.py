#!/usr/bin/env python3
"""Per-phase SM-cycle breakdown of the K3 recurrence kernel (cfg3 geometry): clock64 counters of thread 0 of stream 0
in every CTA (NnamRnnDesc.debug_cycles).  usage: gpu_profile_rnn.py [slots per stream: 16|32|64] [bf16|fp32] [lstm|gru]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nnacousticmodeling_b200 as nn
from nnacousticmodeling_b200 import recurrent_engine as R
from nnacousticmodeling_b200 import synth

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
net = sys.argv[3] if len(sys.argv) > 3 else "lstm"
n_utt = int(sys.argv[4]) if len(sys.argv) > 4 else 1344
x, off, _ = synth.synth_set(1234, n_utt)
m = nn.get_nn(net, 4, [512], 1909, nn.F.relu, [5]); m.init_params(40, np.random.default_rng(1)); m.precision = prec
dev = torch.device("cuda:0")
xd = torch.from_numpy(x).to(dev); out = torch.empty((len(x), 1909), device=dev)
for _ in range(2):
    R.forward_utterances(m, xd, off, out, 0, len(off) - 1, timedelay=5, device=0, nb=nb)
R.PROFILE_CYCLES = []
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); R.forward_utterances(m, xd, off, out, 0, len(off) - 1, timedelay=5, device=0, nb=nb); e.record()
torch.cuda.synchronize()
print(f"{net} nb={nb} {prec}: total {s.elapsed_time(e):.1f} ms")
names = ["gx_issue", "wait_group", "h_load", "mma", "tmem_ld", "gates+store", "fence+publish", "steps"]
for l, buf in enumerate(R.PROFILE_CYCLES):
    a = buf.cpu().numpy().reshape(-1, 8)
    a = a[a[:, 7] > 0]
    steps = a[:, 7]
    per = a[:, :7].sum(axis=0) / steps.sum()
    if l in (0, len(R.PROFILE_CYCLES) - 1):
        print(f"layer {l}: CTAs {len(a)}, stream-0 steps/CTA min {steps.min()} max {steps.max()}")
        print("   cycles/step: " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, per)) + f"  sum={per.sum():.0f}")
