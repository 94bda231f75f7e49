"""Batch sharding and scalar-statistics reduction for multi-GPU runs.

The loss has no data-path collective: every frame triplet is independent and the loss is a mean
over the local batch (/root/reference/algos/depth.py:325), so the global batch is split
contiguously by rank the way ``DistributedSampler`` does for the reference
(/root/reference/misc/train_utils.py:142-148) and each rank runs the kernels on its shard.  The
only traffic is the optional all-reduce of a handful of scalars for logging -- the reference's
``all_reduce_losses`` (/root/reference/misc/utils.py:10-24, commented out in its training loop) --
plus, outside this package, DDP's gradient all-reduce of the surrounding networks.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the contiguous shard of ``rank``; shards differ by at most one sample."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(global_batch, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def global_mean_from_shards(local_mean: torch.Tensor, local_count: int) -> torch.Tensor:
    """Mean over the global batch from per-rank means over shards of possibly different size:
    one all-reduce of two scalars (sum of mean*count, sum of count)."""
    packed = torch.stack((local_mean.detach().to(torch.float64) * local_count,
                          torch.tensor(float(local_count), dtype=torch.float64, device=local_mean.device)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed)
    return (packed[0] / packed[1]).to(local_mean.dtype)


def reduce_loss_dict(losses: Dict[str, torch.Tensor], local_count: int) -> Dict[str, torch.Tensor]:
    """Global-batch means of a dict of scalar losses with a single coalesced all-reduce."""
    keys = sorted(losses)
    if not keys:
        return {}
    dev = losses[keys[0]].device
    packed = torch.stack([losses[k].detach().to(torch.float64) * local_count for k in keys] +
                         [torch.tensor(float(local_count), dtype=torch.float64, device=dev)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed)
    return {k: (packed[i] / packed[-1]).to(losses[k].dtype) for i, k in enumerate(keys)}
