#!/usr/bin/env python3
"""K2 in isolation: every GEMM shape of the BASELINE configs, timed back to back with CUDA events.
usage: [NNAM_GEMM_2SM=0|1] gpu_gemm_bench.py [iters]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnacousticmodeling_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
M = 65536
shapes = [  # (N, K, out_kind, label)
    (2048, 540, ops.OUT_BF16, "cfg2 L0 540->2048"), (2048, 2048, ops.OUT_BF16, "cfg2 L1-5 2048->2048"),
    (1909, 2048, ops.OUT_F32, "cfg2 out 2048->1909 f32"), (1024, 440, ops.OUT_BF16, "cfg1 L0 440->1024"),
    (1024, 1024, ops.OUT_BF16, "cfg1 L1-5 1024->1024"), (1909, 1024, ops.OUT_F32, "cfg1 out 1024->1909 f32"),
    (2048, 512, ops.OUT_F32, "cfg3 upward 512->2048 f32"), (4096, 1024, ops.OUT_F32, "cfg4 upward 1024->4096 f32"),
]
g = torch.Generator(device=dev).manual_seed(1)
print(f"NNAM_GEMM_2SM={os.environ.get('NNAM_GEMM_2SM', '(default 1)')}  M={M}  iters={iters}")
for n, k, kind, label in shapes:
    ld = ops.round_up(k, 8)
    a = (torch.randn((M, ld), generator=g, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn((n, ld), generator=g, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(n, generator=g, device=dev)
    out = torch.empty((M, ops.round_up(n, 16)), dtype=torch.float32 if kind == ops.OUT_F32 else torch.bfloat16, device=dev)
    for _ in range(5):
        ops.linear_bias_act(a, None, w, None, b, M, n, k, act="relu", out_kind=kind, out=(out, None))
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ops.linear_bias_act(a, None, w, None, b, M, n, k, act="relu", out_kind=kind, out=(out, None))
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"{label:30s} {ms * 1e3:8.1f} us  {2.0 * M * n * k / ms / 1e9:8.1f} TFLOP/s")
