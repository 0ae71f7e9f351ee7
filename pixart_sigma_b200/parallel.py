"""Multi-GPU host logic: batch-sharded inference (independent replicas).

Images are independent, so rank r denoises images [lo, hi) with its own full weight replica and the whole sampling
loop runs with NO collective; the only communication is an optional gather of the final latents.  CFG pairs stay on
one rank (the sampler concatenates [uncond, cond] per image).  One process per GPU, `torch.distributed` for plumbing
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: Optional[int] = None, world: Optional[int] = None):
    """Slice every (n_items, ...) tensor to this rank's images; None entries pass through."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    n = next(t.shape[0] for t in tensors if t is not None)
    lo, hi = shard_bounds(n, rank, world)
    return [None if t is None else t[lo:hi] for t in tensors]


def gather_batch(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Inverse of shard_batch for results: all ranks receive the full (n_items, ...) tensor in image order.
    Shards may be ragged (n_items % world != 0): they are padded to the largest shard for the all_gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((cap,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def rank_seed(base_seed: int, image_index: int) -> int:
    """Per-image noise seed that does not depend on how images are sharded, so an N-GPU run reproduces the 1-GPU run."""
    return (base_seed * 1_000_003 + image_index) % (2 ** 31 - 1)
