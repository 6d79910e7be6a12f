#!/usr/bin/env python3
"""Host-side profile (cProfile) of one end-to-end predict() pass: where the time of the e2e leg goes outside the GPU.
usage: gpu_profile_e2e_host.py [workload] [transfer f16|f32|auto]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
t = sys.argv[2] if len(sys.argv) > 2 else "auto"
if t != "auto":
    os.environ["NNAM_TRANSFER"] = t
b = bench.Bench(w, "fp16", 0, 0)
for _ in range(3):
    b.step_e2e()
torch.cuda.synchronize()
t0 = time.perf_counter(); b.step_e2e(); torch.cuda.synchronize(); print("pass ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable(); b.step_e2e(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
