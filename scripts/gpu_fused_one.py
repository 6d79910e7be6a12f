#!/usr/bin/env python3
"""One launch of the fused output-layer kernel next to the unfused pair, timed with CUDA events (and the target of
`ncu --set full -k regex:gemm_logsoftmax`).  usage: gpu_fused_one.py [M] [K] [compact:0|1] [scatter:0|1]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnacousticmodeling_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
compact = len(sys.argv) > 3 and sys.argv[3] == "1"
scatter = len(sys.argv) > 4 and sys.argv[4] == "1"
N = 1909
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
a = torch.randn((M, K), device=dev, generator=g).to(torch.float16)
w = (torch.randn((N, K), device=dev, generator=g) / K ** 0.5).to(torch.float16)
b = torch.randn(N, device=dev, generator=g)
rm = torch.randperm(M, device=dev, generator=g).to(torch.int32) if scatter else None
out = torch.empty((M, N), device=dev)
o16 = torch.empty((M, 1912), dtype=torch.float16, device=dev); ref = torch.empty(M, device=dev)
logits = torch.empty((M, 1920), device=dev)
def fused():
    if compact: ops.linear_logsoftmax(a, None, w, None, b, M, N, K, elem=ops.ELEM_F16, out16=(o16, ref), out_row_map=rm)
    else: ops.linear_logsoftmax(a, None, w, None, b, M, N, K, elem=ops.ELEM_F16, out=out, out_row_map=rm)
def unfused():
    ops.linear_bias_act(a, None, w, None, b, M, N, K, out_kind=ops.OUT_F32, elem=ops.ELEM_F16, out=(logits, None))
    if compact: ops.head(logits, N, rows=M, out16=(o16, ref), out_row_map=rm)
    else: ops.head(logits, N, rows=M, out=out, out_row_map=rm)
for name, fn in (("fused", fused), ("unfused", unfused)):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    print(f"{name}: M={M} K={K} compact={compact} scatter={scatter}: {s.elapsed_time(e) / 10 * 1e3:.1f} us")
