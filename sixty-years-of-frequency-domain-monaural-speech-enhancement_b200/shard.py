"""Data-parallel sharding of a batch of utterances over ranks.

Utterances are independent end to end (per-clip RMS constant, eval-mode BatchNorm, no
cross-clip statistic), so the path shards with NO data-path collective: rank r enhances the
contiguous block of clips ``shard_range(B, r, W)``.  The only exchange is the final gather of
enhanced waveforms, one ``all_gather`` of [B/W, N] fp32 per batch (NCCL over NVLink on the GPU
box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int):
    """Contiguous block split; the first ``batch % world`` ranks take one extra clip."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_waveforms(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """local [b_r, N] -> [batch, N] on every rank, clips in their original order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    n = local.shape[1]
    sizes = [shard_range(batch, r, world) for r in range(world)]
    maxb = max(e - s for s, e in sizes)
    if all(e - s == maxb for s, e in sizes):
        out = torch.empty(batch, n, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros(maxb, n, device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:e - s] for r, (s, e) in enumerate(sizes)], dim=0)
