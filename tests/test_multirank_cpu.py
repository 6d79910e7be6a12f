"""N > 1 host-side logic on CPU: 2 gloo ranks (torch.distributed.run, 127.0.0.1), and bench.py's reference arm under
torchrun (rank 0 prints the one JSON line, the other ranks exit 0 without work)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_two_gloo_ranks_shard_like_the_gpu_ranks():
    res = _torchrun([os.path.join(ROOT, "tests", "_gloo_worker.py")])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")]
    assert len(lines) == 1
    r = json.loads(lines[0][len("RESULT "):])
    assert r["ff_equal"] is True
    assert r["ff_cover"][0][0] == 0 and r["ff_cover"][0][1] == r["ff_cover"][1][0] and r["ff_cover"][1][1] == r["n"]
    assert r["rnn_cover"][0][0] == 0 and r["rnn_cover"][0][1] == r["rnn_cover"][1][0] and r["rnn_cover"][1][1] == r["n"]
    assert r["rnn_max_abs_diff"] < 1e-5
    assert r["max_ms"] == 11.0


def test_bench_reference_arm_under_torchrun_prints_one_line():
    res = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup",
                     "0", "--workload", "cfg1", "--cpu-sample", "2048"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "frames/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["value"] > 0


def test_partition_helpers_cover_and_balance():
    import numpy as np
    sys.path.insert(0, ROOT)
    from nnacousticmodeling_b200 import dist_util
    from oracle import nnam_oracle as O
    _, offsets, _ = O.synth_set(9, 200)
    n = int(offsets[-1])
    for world in (1, 2, 4, 8):
        cuts = [dist_util.rank_shard(offsets, n, True, world, r) for r in range(world)]
        assert cuts[0][2] == 0 and cuts[-1][3] == n
        assert all(cuts[i][3] == cuts[i + 1][2] for i in range(world - 1))
        sizes = np.array([c[3] - c[2] for c in cuts])
        assert sizes.max() - sizes.min() <= 2 * 780  # within two maximum-length utterances
        fcuts = [dist_util.rank_shard(offsets, n, False, world, r) for r in range(world)]
        assert fcuts[0][2] == 0 and fcuts[-1][3] == n and all(fcuts[i][3] == fcuts[i + 1][2] for i in range(world - 1))
    assert dist_util.halo_range(0, 10, 5, 100) == (0, 15) and dist_util.halo_range(90, 100, 5, 100) == (85, 100)
