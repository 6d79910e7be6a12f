"""Process-per-GPU plumbing for the benchmark / multi-process callers (the path shards by independent
utterance / frame ranges, so there is NO data-path collective -- torch.distributed is used only to line the ranks
up (barrier) and to take the max of the per-rank device times)."""
from __future__ import annotations

import os

import numpy as np


def env_world():
    """(world_size, rank, local_rank) from the torchrun environment; (1, 0, 0) when launched plainly."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend, device=None):
    """Join the process group described by the environment (no-op for a single process).  Returns world size."""
    world, _, _ = env_world()
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
            dist.init_process_group(backend, **kw)
    return world


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe root), so that
    the pinned host buffers it allocates next are placed next to the GPU that will DMA into them.  Returns the CPU
    list used, or None when NVML or the container's cpuset does not allow it (then nothing changes)."""
    try:
        import pynvml as nv
        import torch
        nv.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:  # noqa: BLE001
            h = nv.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (max(n_cpu, 1024) + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001
        return None


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """MAX all-reduce of one float (per-rank device time -> job time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def halo_range(f0, f1, splice, n_total):
    """Raw-frame rows a shard [f0, f1) of a feed-forward net must see: +-splice frames from its neighbours, clamped
    at the ends of the WHOLE array only (reference quirk Q1: kw_nn_utils.py:27-33)."""
    return max(int(f0) - int(splice), 0), min(int(f1) + int(splice), int(n_total))


def rank_shard(offsets, n_frames, recurrent, world, rank):
    """The slice of one data set that ``rank`` of ``world`` processes: (u0, u1, f0, f1).  Recurrent nets are cut at
    utterance boundaries balanced on frame counts, feed-forward nets at equal frame counts."""
    from .engine import partition_frames, partition_utterances
    if recurrent:
        offsets = np.asarray(offsets)
        u0, u1 = partition_utterances(offsets, world)[rank]
        return u0, u1, int(offsets[u0]), int(offsets[u1])
    f0, f1 = partition_frames(n_frames, world)[rank]
    return None, None, f0, f1
