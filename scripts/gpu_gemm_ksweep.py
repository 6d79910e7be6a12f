#!/usr/bin/env python3
"""K2: time vs K at fixed M x N (per-tile fixed cost = intercept).  usage: gpu_gemm_ksweep.py [bias 0|1]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnacousticmodeling_b200 import ops
use_bias = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
dev = torch.device("cuda:0")
M, N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536, 2048
for kind, name in ((ops.OUT_BF16, "bf16"), (ops.OUT_F32, "f32")):
    for k in (64, 128, 256, 512, 768, 1024, 2048):
        a = torch.randn((M, k), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, k), device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(N, device=dev) if use_bias else None
        out = torch.empty((M, N), dtype=torch.float32 if kind == ops.OUT_F32 else torch.bfloat16, device=dev)
        for _ in range(3):
            ops.linear_bias_act(a, None, w, None, b, M, N, k, act="relu", out_kind=kind, out=(out, None))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.linear_bias_act(a, None, w, None, b, M, N, k, act="relu", out_kind=kind, out=(out, None))
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"out {name} bias={int(use_bias)} K={k:5d}: {ms*1e3:8.1f} us  {2.0*M*N*k/ms/1e9:8.1f} TFLOP/s  out-stream {out.numel()*out.element_size()/ms/1e6:7.1f} GB/s")
