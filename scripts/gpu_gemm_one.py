#!/usr/bin/env python3
"""One K2 shape, a few launches (for ncu captures).  usage: gpu_gemm_one.py M N K [bf16|f32]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnacousticmodeling_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
kind = ops.OUT_F32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else ops.OUT_BF16
dev = torch.device("cuda:0")
a = torch.randn((M, K), device=dev).to(torch.bfloat16)
w = (torch.randn((N, K), device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(N, device=dev)
out = torch.empty((M, N), dtype=torch.float32 if kind == ops.OUT_F32 else torch.bfloat16, device=dev)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 3:
        s.record()
    ops.linear_bias_act(a, None, w, None, b, M, N, K, act="relu", out_kind=kind, out=(out, None))
e.record(); torch.cuda.synchronize()
print(f"{M}x{N}x{K}: {s.elapsed_time(e) / 3 * 1e3:.1f} us per launch")
