#!/usr/bin/env python3
"""Bring-up diagnostics for K1/K2/K4 on a real B200 (prints error structure, not just pass/fail)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnacousticmodeling_b200 import ops  # noqa: E402
from oracle import nnam_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)


def bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


def gemm_case(M, N, K, nsplit, out_kind, act="relu", seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    ad, wd, bd = (torch.from_numpy(t).to(dev) for t in (a, w, b))
    kind = ops.OUT_BF16_SPLIT if nsplit == 3 else ops.OUT_BF16
    a_hi, a_lo = ops.convert_f32(ad, kind)
    w_hi, w_lo = ops.convert_f32(wd, kind)
    hi, lo = ops.linear_bias_act(a_hi, a_lo, w_hi, w_lo, bd, M, N, K, act=act, out_kind=out_kind, nsplit=nsplit)
    torch.cuda.synchronize()
    got = hi.float()
    if lo is not None:
        got = got + lo.float()
    got = got[:, :N].cpu().numpy()
    if nsplit == 1:
        ref = bf16_round(a).astype(np.float64) @ bf16_round(w).astype(np.float64).T + b
    else:
        ref = a.astype(np.float64) @ w.astype(np.float64).T + b
    ref = O.activation(act)(ref)
    err = np.abs(got - ref)
    tol = {ops.OUT_F32: 2e-5 if nsplit == 1 else 1e-4, ops.OUT_BF16: 2e-2, ops.OUT_BF16_SPLIT: 1e-4}[out_kind]
    ok = err.max() < tol * max(1.0, np.abs(ref).max())
    print(f"gemm M={M} N={N} K={K} nsplit={nsplit} out={out_kind} act={act}: max_err={err.max():.3e} "
          f"ref_max={np.abs(ref).max():.2f} {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        bad = err > tol * max(1.0, np.abs(ref).max())
        rows = np.where(bad.any(axis=1))[0]
        cols = np.where(bad.any(axis=0))[0]
        print("   bad rows:", rows[:20], "... n=", len(rows), " bad cols:", cols[:20], "... n=", len(cols))
        print("   got[0,:8]", got[0, :8], "\n   ref[0,:8]", ref[0, :8])
    return ok


def main():
    ok = True
    # ---- K1
    g = np.load("tests/golden/splice.npz")
    ft = O.load_kaldi_feature_transform("tests/golden/final.feature_transform")
    x = torch.from_numpy(g["x"]).to(dev)
    add, mul = torch.from_numpy(ft["addShift"]).to(dev), torch.from_numpy(ft["rescale"]).to(dev)
    out, _ = ops.splice_transform(x, x.shape[0], 5, add, mul)
    e = np.array_equal(out.cpu().numpy(), g["splice11_ft"])
    print("K1 golden bit-exact:", e, flush=True)
    ok &= e
    out, _ = ops.splice_transform(x, x.shape[0], 5)
    e = np.array_equal(out.cpu().numpy(), g["splice11"])
    print("K1 golden (no ft) bit-exact:", e, flush=True)
    ok &= e
    # ---- K4
    h = np.load("tests/golden/head.npz")
    ap = torch.from_numpy(np.load("tests/golden/log_ap_Kaldi1909.npy")).to(dev)
    y = torch.from_numpy(h["y"]).to(dev)
    o1 = ops.head(y, 1909).cpu().numpy()
    o2 = ops.head(y, 1909, prior=ap.reshape(-1)).cpu().numpy()
    e1, e2 = np.abs(o1 - h["logsoftmax"]).max(), np.abs(o2 - h["head_ap"]).max()
    print(f"K4 max err logsoftmax={e1:.3e} with prior={e2:.3e}", flush=True)
    ok &= bool(e1 < 1e-3 and e2 < 1e-3)
    # ---- K2
    for (M, N, K) in [(128, 256, 64), (128, 16, 64), (256, 512, 128), (300, 1024, 544), (1000, 1909, 440),
                      (4096, 2048, 2048), (77, 40, 40)]:
        for nsplit, outk in [(1, ops.OUT_F32), (1, ops.OUT_BF16), (3, ops.OUT_F32), (3, ops.OUT_BF16_SPLIT)]:
            ok &= gemm_case(M, N, K, nsplit, outk)
    ok &= gemm_case(512, 1024, 1024, 3, ops.OUT_F32, act="sigmoid")
    ok &= gemm_case(512, 1024, 1024, 1, ops.OUT_F32, act="tanh")
    ok &= gemm_case(512, 1024, 1024, 1, ops.OUT_F32, act="identity")
    # ---- quick timing of the big GEMM (bf16), 20 launches
    M, N, K = 131072, 2048, 2048
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    b = torch.zeros(N, device=dev)
    out = (torch.empty(M, N, device=dev, dtype=torch.bfloat16), None)
    for _ in range(3):
        ops.linear_bias_act(a, None, w, None, b, M, N, K, act="relu", out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        ops.linear_bias_act(a, None, w, None, b, M, N, K, act="relu", out=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"gemm {M}x{N}x{K} bf16: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    t0 = time.time()
    ref = torch.relu(a.float()[:4096] @ w.float().T)
    err = (out[0][:4096].float() - ref).abs().max().item()
    print(f"   check vs torch on 4096 rows: max err {err:.3e} (ref max {ref.abs().max().item():.1f})")
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
