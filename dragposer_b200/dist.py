"""Clip sharding across the GPUs of one box.

Clips are independent (no cross-clip term anywhere in python/src/drag_pose.py), so the
clip axis is partitioned contiguously, one process per GPU, with ZERO traffic inside the
optimisation loop; the only collective is the gather of the result rows
(pose (B/G,88) | global_pos (B/G,3)) in rank order, so row c of the gathered tensor is
clip c.  Works over NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ROW = 91  # 88 standardised pose values + 3 root position


def shard_bounds(n_clips: int, world_size: int, rank: int):
    """Contiguous equal shards; the last ranks may own padding rows (n_clips is padded
    up to a multiple of world_size so that the all-gather has equal contributions)."""
    per = (n_clips + world_size - 1) // world_size
    lo = min(rank * per, n_clips)
    hi = min(lo + per, n_clips)
    return lo, hi, per


def pack_rows(pose: torch.Tensor, gpos: torch.Tensor, per: int) -> torch.Tensor:
    """(n,88),(n,3) -> (per,91) with zero padding rows."""
    n = pose.shape[0]
    rows = torch.zeros((per, ROW), dtype=torch.float32, device=pose.device)
    rows[:n, :88] = pose
    rows[:n, 88:] = gpos
    return rows


def gather_results(pose: torch.Tensor, gpos: torch.Tensor, n_clips: int, group=None, out: torch.Tensor = None):
    """All-gather the local result rows in rank order -> pose (n_clips,88), gpos (n_clips,3)
    on every rank.  `out` (world*per, 91) may be preallocated to avoid per-frame allocations."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi, per = shard_bounds(n_clips, world, rank)
    assert pose.shape[0] == hi - lo, "local shard size does not match shard_bounds"
    rows = pack_rows(pose, gpos, per)
    if world == 1:
        full = rows
    else:
        full = out if out is not None else torch.empty((world * per, ROW), dtype=torch.float32, device=pose.device)
        dist.all_gather_into_tensor(full, rows, group=group)
    return full[:n_clips, :88], full[:n_clips, 88:]


def gather_frames(pose: torch.Tensor, gpos: torch.Tensor, n_clips: int, group=None):
    """A whole batch of frames in ONE collective ("or once per run for all T frames"):
    local pose (T,n,88), gpos (T,n,3) -> pose (T,n_clips,88), gpos (T,n_clips,3) on every rank, row c = clip c.
    Shards are the contiguous equal shards of shard_bounds (the last one may be short; it is zero padded on the wire)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi, per = shard_bounds(n_clips, world, rank)
    T, n = pose.shape[0], pose.shape[1]
    assert n == hi - lo, "local shard size does not match shard_bounds"
    rows = torch.zeros((T, per, ROW), dtype=torch.float32, device=pose.device)
    rows[:, :n, :88] = pose
    rows[:, :n, 88:] = gpos
    if world == 1:
        full = rows
    else:
        wire = torch.empty((world * T, per, ROW), dtype=torch.float32, device=pose.device)  # rank-major concatenation
        dist.all_gather_into_tensor(wire, rows, group=group)
        full = wire.view(world, T, per, ROW).permute(1, 0, 2, 3).reshape(T, world * per, ROW)  # clip order inside every frame
    return full[:, :n_clips, :88], full[:, :n_clips, 88:]
