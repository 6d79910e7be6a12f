#!/usr/bin/env python3
"""K1 in isolation: splice + transform + i-vector append into the bf16 GEMM input, 65,536-frame chunks of a
1.1 M-frame set, CUDA events over back-to-back launches (inputs larger than L2 across the sweep)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nnacousticmodeling_b200 as nn
from nnacousticmodeling_b200 import ops
dev = torch.device("cuda:0")
n, chunk = 1124823, 65536
x = torch.randn((n, 40), device=dev); iv = torch.randn((n, 100), device=dev)
ft = nn.loadKaldiFeatureTransform(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "final.feature_transform"))
add = torch.from_numpy(ft["addShift"]).to(dev); mul = torch.from_numpy(ft["rescale"]).to(dev)
for kind, name, bpf in ((ops.OUT_BF16, "bf16", 160 + 400 + 1088), (ops.OUT_F32, "f32", 160 + 400 + 2160)):
    ld = 544 if kind == ops.OUT_BF16 else 540
    out = torch.empty((chunk, ld), dtype=torch.bfloat16 if kind == ops.OUT_BF16 else torch.float32, device=dev)
    def sweep():
        for c0 in range(0, n - chunk, chunk):
            ops.splice_transform(x, n, 5, add, mul, iv[c0:c0 + chunk], f0=c0, f1=c0 + chunk, out_kind=kind, ldo=ld, out=(out, None))
    sweep(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); sweep(); sweep(); e.record(); torch.cuda.synchronize()
    launches = 2 * len(range(0, n - chunk, chunk))
    us = s.elapsed_time(e) * 1e3 / launches
    print(f"splice -> {name}: {us:6.1f} us per 65,536 frames, {chunk * bpf / us / 1e3:7.1f} GB/s algorithmic")
