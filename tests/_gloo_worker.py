"""Worker for tests/test_multirank_cpu.py: launched by torch.distributed.run with 2 processes on CPU (gloo).

Each rank takes ITS shard of one synthetic set exactly as the GPU ranks do (dist_util.rank_shard / halo_range),
evaluates it with the CPU oracle (the GPU kernels cannot run here; the oracle stands in for them so that the
HOST-SIDE sharding logic is what is under test), the shards are gathered, and rank 0 checks that the concatenation
is bit-identical to the unsharded result.  Also exercises barrier + MAX-over-ranks timing."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nnacousticmodeling_b200 import dist_util  # noqa: E402
from oracle import nnam_oracle as O  # noqa: E402


def main():
    world = dist_util.init("gloo")
    _, rank, _ = dist_util.env_world()
    assert world == 2
    x, offsets, _ = O.synth_set(5, 9)
    n = len(x)
    res = {}
    # ---- feed-forward: frame shards with a +-splice halo (quirk Q1 stays bit-exact across shard borders)
    splice = 5
    p = O.init_mlp(np.random.default_rng(3), 40 * (2 * splice + 1), 32, 2, 7)
    full = O.log_softmax(O.mlp_forward(p, O.splicing(x, range(-splice, splice + 1)), 2))
    _, _, f0, f1 = dist_util.rank_shard(offsets, n, False, world, rank)
    lo, hi = dist_util.halo_range(f0, f1, splice, n)
    # local splice over the halo window; frames whose context leaves the window are only the global ends
    idx = np.clip(np.arange(f0, f1)[:, None] + np.arange(-splice, splice + 1)[None, :], 0, n - 1) - lo
    assert idx.min() >= 0 and idx.max() < hi - lo
    feats = x[lo:hi][idx].reshape(f1 - f0, -1)
    mine = O.log_softmax(O.mlp_forward(p, feats, 2))
    parts = [None] * world
    dist.all_gather_object(parts, (f0, f1, mine))
    if rank == 0:
        cat = np.concatenate([m for _, _, m in sorted(parts, key=lambda t: t[0])])
        res["ff_equal"] = bool(np.array_equal(cat, full))
        res["ff_cover"] = [[int(a), int(b)] for a, b, _ in sorted(parts, key=lambda t: t[0])]
    # ---- recurrent: utterance shards balanced on frames
    pr = O.init_recurrent(np.random.default_rng(4), "lstm", 40, 16, 1, 7)
    full_r = O.predict(O.RecurrentNet(pr, "lstm", 1), x, offsets, "lstm", 1, 2, None)
    u0, u1, g0, g1 = dist_util.rank_shard(offsets, n, True, world, rank)
    mine_r = O.predict(O.RecurrentNet(pr, "lstm", 1), x[g0:g1], offsets[u0:u1 + 1] - offsets[u0], "lstm", 1, 2, None)
    dist.all_gather_object(parts, (g0, g1, mine_r))
    if rank == 0:
        cat = np.concatenate([m for _, _, m in sorted(parts, key=lambda t: t[0])])
        res["rnn_max_abs_diff"] = float(np.abs(cat - full_r).max())  # batch composition changes BLAS blocking only
        res["rnn_cover"] = [[int(a), int(b)] for a, b, _ in sorted(parts, key=lambda t: t[0])]
    # ---- timing plumbing: barrier, MAX over ranks
    dist_util.barrier()
    res["max_ms"] = dist_util.max_over_ranks(10.0 + rank)
    if rank == 0:
        res["n"] = n
        print("RESULT " + json.dumps(res))
    dist_util.finalize()


if __name__ == "__main__":
    main()
