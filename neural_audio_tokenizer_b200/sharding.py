"""Frame / clip sharding across the GPUs of one box and the all-gather of the index streams.

Frames are independent inside the RVQ (the only sequential dependency is across layers within a frame,
nat.py:1398-1405), so the path shards with no data-path collective: codebooks are replicated, every rank encodes a
contiguous frame range, and only the int16 index streams ([L_total, frames] = 16 bytes per frame for 8 layers) are
exchanged. NCCL has no 16-bit integer type, so the payload travels as bytes. Works on any torch.distributed backend
(`nccl` on the B200 box, `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` for `rank`: equal ceil-sized shards, the tail ranks may be short/empty."""
    per = -(-n_items // world_size) if world_size > 0 else n_items
    start = min(n_items, rank * per)
    return start, min(n_items, start + per)


def shard_clips(n_clips: int, world_size: int, rank: int):
    """Round-robin clip ids for a corpus (BASELINE.json config 3: clip_id mod G)."""
    return range(rank, n_clips, world_size)


def all_gather_codes(local_codes: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local_codes [L, n_local] (any integer dtype, K <= 32768) for this rank's `shard_range` -> [L, n_total] int16
    on every rank. Shards are padded to the common ceil size because all-gather needs equal counts."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    L = local_codes.shape[0]
    per = -(-n_total // world)
    start, stop = shard_range(n_total, world, rank)
    if local_codes.shape[1] != stop - start:
        raise ValueError(f"rank {rank} holds {local_codes.shape[1]} frames, its shard is {stop - start}")
    send = torch.zeros((L, per), dtype=torch.int16, device=local_codes.device)
    send[:, :stop - start] = local_codes.to(torch.int16)
    if world == 1:
        return send[:, :n_total].contiguous()
    recv = torch.empty((world, L, per), dtype=torch.int16, device=local_codes.device)
    dist.all_gather_into_tensor(recv.view(torch.uint8).view(-1), send.view(torch.uint8).view(-1), group=group)
    # [world, L, per] -> [L, world * per] -> drop the padding of the last shards
    return recv.permute(1, 0, 2).reshape(L, world * per)[:, :n_total].contiguous()
