"""Frame sharding across the GPUs of one node.

Blocks - and therefore frames - are classified independently, every rank holds a replica of the four
stage networks, so the inference path needs NO collective: rank r runs the cascade on a contiguous
range of frames.  The only communication is the final gather of uint8 labels (32,400 B per 4K frame) to
rank 0, done with torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Two forms of the gather:
* `gather_labels`       - one gather of every rank's whole label vector after its last cascade;
* `ChunkedLabelGather`  - one asynchronous gather per cascade chunk, issued on the chunk's own stream right after
                          its cascade, so the transfer of chunk c runs behind the compute of chunk c + 1 and only
                          the last chunk's (small) gather is exposed.  This is what the strong-scaling bench uses:
                          with 8 frames per GPU a step is ~17 ms and a trailing 2 MB gather is no longer noise.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_frames(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced frame range of `rank`: (first_frame, frame_count).  Earlier ranks get the remainder."""
    if not 0 <= rank < world_size:
        raise ValueError("rank outside the world")
    base, rem = divmod(n_frames, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_labels(local: torch.Tensor, n_frames: int, blocks_per_frame: int, rank: int, world_size: int,
                  group=None) -> Optional[torch.Tensor]:
    """Gather per-rank label vectors (frame order) on rank 0.  Returns the full vector on rank 0, None elsewhere."""
    first, count = shard_frames(n_frames, rank, world_size)
    if local.numel() != count * blocks_per_frame:
        raise ValueError(f"rank {rank} holds {local.numel()} labels, expected {count * blocks_per_frame}")
    if world_size == 1:
        return local
    max_count = shard_frames(n_frames, 0, world_size)[1]
    padded = torch.zeros(max_count * blocks_per_frame, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world_size)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0, group=group)
    if rank != 0:
        return None
    parts = [bufs[r][: shard_frames(n_frames, r, world_size)[1] * blocks_per_frame] for r in range(world_size)]
    return torch.cat(parts)


class ChunkedLabelGather:
    """Per-chunk asynchronous gather of a sharded label vector onto rank 0.

    Every rank owns `local` (uint8 [max_count * blocks_per_frame], max_count = the largest shard; ranks with a shorter
    shard leave the tail unused).  `gather_chunk(f0, nf)` sends the labels of local frames [f0, f0 + nf) - call it on
    the stream that produced them; it returns immediately.  `finish()` waits for all chunks and, on rank 0, returns the
    full vector in global frame order (a view-assembled copy; None elsewhere).  Buffers are allocated once and reused
    across steps."""

    def __init__(self, n_frames: int, blocks_per_frame: int, rank: int, world_size: int, device, dtype=torch.uint8, group=None):
        self.n_frames, self.bpf, self.rank, self.world, self.group = n_frames, blocks_per_frame, rank, world_size, group
        self.first, self.count = shard_frames(n_frames, rank, world_size)
        self.max_count = shard_frames(n_frames, 0, world_size)[1]
        self.local = torch.zeros(self.max_count * blocks_per_frame, dtype=dtype, device=device)
        # rank 0: one landing buffer per rank, same padded shape, so every chunk gathers equal-sized views
        self.landing: Optional[List[torch.Tensor]] = None
        if rank == 0 and world_size > 1:
            self.landing = [torch.zeros_like(self.local) for _ in range(world_size)]
        self._pending = []

    def gather_chunk(self, f0: int, nf: int) -> None:
        if self.world == 1 or nf <= 0:
            return
        lo, hi = f0 * self.bpf, (f0 + nf) * self.bpf
        views = [b[lo:hi] for b in self.landing] if self.rank == 0 else None
        self._pending.append(dist.gather(self.local[lo:hi], views, dst=0, group=self.group, async_op=True))

    def finish(self) -> Optional[torch.Tensor]:
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self.rank != 0:
            return None
        if self.world == 1:
            return self.local[: self.count * self.bpf]
        parts = [self.landing[r][: shard_frames(self.n_frames, r, self.world)[1] * self.bpf] for r in range(self.world)]
        return torch.cat(parts)
