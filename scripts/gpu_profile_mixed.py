#!/usr/bin/env python3
"""Timing of the two concurrent recurrence launches of a MixedSchedule (cfg3 geometry).
usage: gpu_profile_mixed.py [n_utt] [long groups: 0 = plain schedule, -1 = cost model, 1-6] [lstm|gru|blstm|bgru] [long batches]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 3696
g_a = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = sys.argv[3] if len(sys.argv) > 3 else "lstm"
if g_a == 0:
    os.environ["NNAM_RNN_MIXED"] = "0"
elif g_a < 0:
    pass  # the cost model decides
else:
    os.environ["NNAM_RNN_MIXED"] = "force"
    os.environ["NNAM_RNN_MIXED_GROUPS"] = str(g_a)
    if len(sys.argv) > 4:
        os.environ["NNAM_RNN_MIXED_BATCHES"] = sys.argv[4]
import nnacousticmodeling_b200 as nn
from nnacousticmodeling_b200 import recurrent_engine as R, ops
from nnacousticmodeling_b200 import synth
x, off, _ = synth.synth_set(1234, n_utt)
bid = net in ("blstm", "bgru")
m = nn.get_nn(net, 4, [512], 1909, nn.F.relu, [5]); m.init_params(40, np.random.default_rng(1)); m.precision = os.environ.get("PREC", "fp16")
td = 0 if bid else 5
dev = torch.device("cuda:0")
xd = torch.from_numpy(x).to(dev); out = torch.empty((len(x), 1909), device=dev)
for _ in range(2):
    R.forward_utterances(m, xd, off, out, 0, len(off) - 1, timedelay=td, device=0)
marks = []
orig = ops.rnn_seq
def timed(desc, flops):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig(desc, flops); e.record()
    marks.append((desc.batch, desc.n_groups, s, e))
ops.rnn_seq = timed
R.ops.rnn_seq = timed
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); R.forward_utterances(m, xd, off, out, 0, len(off) - 1, timedelay=td, device=0); e.record()
torch.cuda.synchronize()
plan = next(iter(m._plans.values()))
print(f"n_utt={n_utt} long groups={g_a}: pass {s.elapsed_time(e):.2f} ms, schedule {[v[1] for v in plan._sched_cache.values()]}")
for b, g, a, z in marks:
    print(f"   K3 launch: {b:3d} slots x {g} groups: {a.elapsed_time(z):.2f} ms")
