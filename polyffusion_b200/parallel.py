"""Multi-GPU plumbing for the sampling path: one process per GPU, batch sharded by contiguous
blocks, no collective inside a step, a single all-gather of the finished piano-rolls at the end
(SURVEY.md section 8e).  Works with any torch.distributed backend (NCCL over NVLink on B200 boxes;
gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` samples owned by `rank`; earlier ranks take the remainder."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], world_size, rank)
    return t[lo:hi]


def sample_seed(base_seed: int, global_sample_index: int) -> int:
    """Noise-stream seed tied to the GLOBAL sample index, so results do not depend on the rank count."""
    return (base_seed * 1000003 + global_sample_index) % (2**63 - 1)


def gather_samples(x_local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank sample blocks into the full [total, ...] tensor on every rank.
    Shards may be ragged (total % world != 0): blocks are padded to the largest shard."""
    if not dist.is_available() or not dist.is_initialized():
        return x_local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = x_local
    if x_local.shape[0] < max_n:
        pad = torch.cat([x_local, x_local.new_zeros((max_n - x_local.shape[0], *x_local.shape[1:]))])
    outs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad.contiguous(), group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)])


def sample_sharded(sampler_fn, cond: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None):
    """Run `sampler_fn(cond_shard, lo, hi) -> samples` on this rank's shard and gather the result."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(total, world, rank)
    local = sampler_fn(cond[lo:hi], lo, hi)
    return gather_samples(local, total, group)
