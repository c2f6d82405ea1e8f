"""Frame sharding across the GPUs of one box: frames are independent, so ranks own contiguous
frame blocks and the only exchange is a gather of the fixed-size keypoint records.  Backend-agnostic
(`nccl` on the GPUs, `gloo` in the CPU tests)."""
from __future__ import annotations

import numpy as np

from .capi import HUMAN_DT


def shard_range(n_frames: int, world: int, rank: int):
    """contiguous block of frames owned by `rank` (GPU g gets frames [g*B, (g+1)*B) of each super-batch)"""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(humans_u8, counts_i32, world: int, group=None):
    """all-gather of padded `hp_human[B][cap]` bytes + `int count[B]` (equal-sized blocks on every rank).
    humans_u8: uint8 tensor [B*cap*292]; counts_i32: int32 tensor [B].  Returns (all_humans, all_counts)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return humans_u8, counts_i32
    gh = torch.empty(world * humans_u8.numel(), dtype=torch.uint8, device=humans_u8.device)
    gc = torch.empty(world * counts_i32.numel(), dtype=torch.int32, device=counts_i32.device)
    dist.all_gather_into_tensor(gh, humans_u8, group=group)
    dist.all_gather_into_tensor(gc, counts_i32, group=group)
    return gh, gc


def unpack_records(humans_u8: np.ndarray, counts: np.ndarray, cap: int):
    """bytes -> list (one per frame) of structured HUMAN_DT arrays"""
    rec = np.frombuffer(np.ascontiguousarray(humans_u8).tobytes(), dtype=HUMAN_DT).reshape(len(counts), cap)
    return [rec[i, :int(counts[i])].copy() for i in range(len(counts))]
