"""Patch sharding across the GPUs of one box (one process per GPU, torch.distributed for the plumbing).

Every patch is an independent diffusion chain: rank r samples patches [lo, hi) of the batch with
its slice of `cond` (and of any injected noise); there is NO collective inside the sampling loop.  The only
exchange is one final all_gather of the finished patches (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n patches: the first n % world ranks get one extra patch."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def gather_patches(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """all_gather of per-rank results back into batch order (ragged shards are padded to the largest shard)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


def sample_sharded(sample_fn: Callable[..., torch.Tensor], cond: torch.Tensor, noise: Optional[Sequence[torch.Tensor]] = None,
                   group=None, diffusion=None) -> torch.Tensor:
    """Run `sample_fn(cond_shard, noise=noise_shards)` on this rank's patches and gather the full batch.
    `noise` (if given) is the GLOBAL list [x_T, n_1, ...]; each rank uses its batch slice.  Without injected noise pass the
    `GaussianDiffusion` that `sample_fn` drives as `diffusion`: its in-kernel Philox generator is then indexed by the GLOBAL
    patch index (`noise_shard = (lo, total)`), so either way results do not depend on the number of ranks -- every rank
    using the same (seed, offset 0) stream would make patch i of every shard share its noise."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    total = cond.shape[0]
    c = shard_batch(cond, rank, world).contiguous()
    nz = [shard_batch(n, rank, world).contiguous() for n in noise] if noise is not None else None
    if c.shape[0] == 0:
        raise ValueError("more ranks than patches: give every rank at least one patch")
    if diffusion is not None:
        diffusion.noise_shard = (shard_range(total, rank, world)[0], total)
    try:
        local = sample_fn(c, noise=nz) if nz is not None else sample_fn(c)
    finally:
        if diffusion is not None:
            diffusion.noise_shard = None
    return gather_patches(local, total, group)
