"""Instance sharding across the GPUs of one box: the batch axis (Monte-Carlo samples, sweep points, AC
frequencies) splits into contiguous blocks, one per rank; every rank solves its block with no data-path collective, and
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only to gather per-instance results at the end.
"""
import numpy as np


def shard_bounds(n_instances, rank, world):
    """[lo, hi) of rank's contiguous block: ceil(n/world) per rank, the last ranks may be short or empty."""
    per = -(-n_instances // world)
    lo = min(rank * per, n_instances)
    return lo, min(lo + per, n_instances)


def gather_instances(local, n_instances, group=None, device=None):
    """All-gather per-instance results. ``local`` is this rank's block ([hi-lo, ...] numpy array); returns the full
    [n_instances, ...] array on every rank. Blocks are padded to the common block size for the collective."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_instances // world)
    local = np.ascontiguousarray(local)
    lo, hi = shard_bounds(n_instances, rank, world)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[: hi - lo] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    full = np.concatenate([o.cpu().numpy() for o in out], axis=0)
    return full[:n_instances]
