"""Multi-GPU plumbing (one process per GPU, torch.distributed): which rank owns which targets and
how the per-rank results are put back together.  No numerics here: the compute callables are the
C-ABI evaluators (GPU) -- tests substitute the oracle to exercise this logic on CPU with gloo.

Direct sum: every rank holds all sources, computes the accelerations of its own
contiguous target shard [ceil(n r/w), ceil(n (r+1)/w)) -- the kd-tree's own split rule -- and the
shards are all-gathered.  There is no reduction: targets are independent."""
import numpy as np

from ._lib import shard_range


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def all_gather_shards(local_shard, n, group=None):
    """local_shard: torch tensor (count_r, 3) of this rank's targets; returns the full (n, 3) tensor"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n, world)
    pad = max(sizes)
    buf = torch.zeros((pad, 3), dtype=local_shard.dtype, device=local_shard.device)
    buf[: local_shard.shape[0]] = local_shard
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def sharded_direct3(compute_shard, pos, n, group=None):
    """compute_shard(rank, world) -> (n, 3) tensor whose rows [begin, end) of this rank are valid
    (what nbco_force_direct3 writes with cfg.rank/cfg.world set); returns the assembled (n, 3)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_range(n, rank, world)
    acc = compute_shard(rank, world)
    return all_gather_shards(acc[b:e].contiguous(), n, group)
