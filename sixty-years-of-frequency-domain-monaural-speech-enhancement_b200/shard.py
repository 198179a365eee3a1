"""Data-parallel sharding of a batch of utterances over ranks.

Utterances are independent end to end (per-clip RMS constant, eval-mode BatchNorm, no
cross-clip statistic), so the path shards with NO data-path collective: rank r enhances the
contiguous block of clips ``shard_range(B, r, W)``.  The only exchange is the final gather of
enhanced waveforms: ``GatherPipeline`` collects every batch on ONE rank (the process that writes the
files) on a side stream, behind the next batch's compute; ``gather_waveforms`` is the blocking all-gather
form (NCCL over NVLink on the GPU box, gloo in the CPU tests).
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


class GatherPipeline:
    """Gather of [b_r, N] enhanced waveforms to rank ``dst``, overlapped with the next batch's compute.

    Round 1 issued ``all_gather_into_tensor`` on the compute stream after every batch: every rank received all
    W x 16.4 MB although only the writer needs them, and the 0.3 ms were serial.  Here the
    collective is a gather to one rank, issued on a side stream that waits for the producing batch only; at most
    ``depth`` gathers are in flight (their buffers stay referenced until the collective has completed)."""

    def __init__(self, batch: int, dst: int = 0, depth: int = 2, group=None):
        self.batch, self.dst, self.depth, self.group = batch, dst, depth, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cuda = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        self.side = torch.cuda.Stream() if self.cuda else None
        self.inflight = []          # (work, local, outs)
        sizes = [shard_range(batch, r, self.world) for r in range(self.world)]
        if len({e - s for s, e in sizes}) != 1:
            raise ValueError("GatherPipeline needs equal shards (pad the batch to a multiple of the world size)")

    def submit(self, local: torch.Tensor):
        """Queue the gather of this rank's ``local`` [b_r, N]; returns the list of per-rank tensors on ``dst`` (valid
        after ``drain`` or after ``depth`` more submits), None elsewhere."""
        local = local.contiguous()
        outs = [torch.empty_like(local) for _ in range(self.world)] if self.rank == self.dst else None
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                work = dist.gather(local, outs, dst=self.dst, group=self.group, async_op=True)
        else:
            work = dist.gather(local, outs, dst=self.dst, group=self.group, async_op=True)
        self.inflight.append((work, local, outs))
        while len(self.inflight) > self.depth:
            self._retire()
        return outs

    def _retire(self):
        work, _local, _outs = self.inflight.pop(0)
        work.wait()          # NCCL: the CURRENT stream waits for the collective; gloo: blocks the host

    def drain(self):
        while self.inflight:
            self._retire()
