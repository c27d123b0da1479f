"""Frame sharding across the GPUs of one node.

Blocks - and therefore frames - are classified independently, every rank holds a replica of the four
stage networks, so the inference path needs NO collective: rank r runs the cascade on a contiguous
range of frames.  The only communication is the final gather of uint8 labels (32,400 B per 4K frame) to
rank 0, done with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

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
