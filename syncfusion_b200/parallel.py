"""Clip-sharded data parallelism for sampling.

Clips are independent (GroupNorm / LayerNorm / attention are all per clip), so the batch is partitioned
contiguously by rank, every rank runs the whole sampling loop locally with zero communication, and the only
collective is ONE all-gather of the finished fp32 waveforms (1 MiB per clip) over NCCL / NVLink.  The reference
samples on a single device (main/generation.py:16,44); this is the new multi-GPU step north_star asks for.
Works with the ``gloo`` backend on CPU tensors too (used by the world_size-2 tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition; the first ``batch % world`` ranks take one extra clip."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(rank: int, world: int, x_noisy: Tensor, channels: Sequence[Tensor], embedding: Tensor):
    lo, hi = shard_bounds(x_noisy.shape[0], world, rank)
    return x_noisy[lo:hi], [c[lo:hi] for c in channels], embedding[lo:hi]


def gather_waveforms(local: Tensor, batch: int, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """All-gather ``[B_local, 1, L]`` waveforms into ``[batch, 1, L]`` on every rank (one collective)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return local
    counts = [shard_bounds(batch, world, r)[1] - shard_bounds(batch, world, r)[0] for r in range(world)]
    cmax = max(counts)
    pad = local
    if local.shape[0] < cmax:
        pad = torch.cat([local, local.new_zeros((cmax - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((world * cmax,) + tuple(local.shape[1:]))
    # Host-side fence on both sides of the collective: the sampling kernels are persistent one-CTA-per-SM grids chained
    # with programmatic dependent launch, and 2-GPU runs in which the NCCL kernel was enqueued straight behind / in front
    # of them showed barrier-wait timeouts on rank 1 (open issue).  One sync per sample() is free
    # next to ~0.4 s of sampling.
    if local.is_cuda:
        torch.cuda.current_stream(local.device).synchronize()
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if local.is_cuda:
        torch.cuda.current_stream(local.device).synchronize()
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax: r * cmax + counts[r]] for r in range(world)], dim=0)


def sample_sharded(sample_fn: Callable[..., Tensor], x_noisy: Tensor, num_steps: int, channels: Sequence[Tensor],
                   embedding: Tensor, embedding_scale: float, group: Optional[dist.ProcessGroup] = None,
                   gather: bool = True) -> Tensor:
    """``sample_fn(x_noisy=, num_steps=, channels=, embedding=, embedding_scale=)`` on this rank's shard + gather."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    xs, cs, es = shard_batch(rank, world, x_noisy, channels, embedding)
    if xs.shape[0] > 0:
        local = sample_fn(x_noisy=xs, num_steps=num_steps, channels=cs, embedding=es, embedding_scale=embedding_scale)
    else:
        local = xs.clone()
    return gather_waveforms(local, x_noisy.shape[0], group) if gather else local
