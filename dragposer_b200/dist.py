"""Clip sharding across the GPUs of one box.

Clips are independent (no cross-clip term anywhere in python/src/drag_pose.py), so the clip axis is partitioned contiguously,
one process per GPU, with ZERO traffic inside the optimisation loop; the only collective is the gather of the result rows in
rank order, so that row c of a gathered frame is clip c.  Works over NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU
tests).

Wire format: packed rows of ROW = 92 floats, [pose 88 | global_pos 3 | pad].  The frame kernels write this layout themselves
(`BatchedDragPose.run_frames_device(..., out_gpos=None)`), so a batch of frames goes on the wire as it lies in HBM -- no zero fill,
no slice copies, no permute.  `gather_packed` sends the shards to the ONE rank that consumes them (round 1 all-gathered every
frame batch to every rank: 8x the bytes anyone read) and can run asynchronously, so the gather of frame batch k overlaps the
kernels of batch k + 1.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ROW = 92  # 88 standardised pose values + 3 root position + 1 pad (keeps every row 16-byte aligned); include/dp_engine.h:DP_ROW


def _world_rank(group=None):
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_bounds(n_clips: int, world_size: int, rank: int):
    """Contiguous equal shards; the last ranks may own padding rows (n_clips is padded
    up to a multiple of world_size so that every rank contributes the same number of rows)."""
    per = (n_clips + world_size - 1) // world_size
    lo = min(rank * per, n_clips)
    hi = min(lo + per, n_clips)
    return lo, hi, per


def pack_rows(pose: torch.Tensor, gpos: torch.Tensor, per: int) -> torch.Tensor:
    """(n,88),(n,3) -> (per,ROW) with zero padding rows (for callers that hold the two arrays separately)."""
    n = pose.shape[0]
    rows = torch.zeros((per, ROW), dtype=torch.float32, device=pose.device)
    rows[:n, :88] = pose
    rows[:n, 88:91] = gpos
    return rows


class GatheredFrames:
    """Result rows of T frames from every rank, as they arrived: buffer (world, T, per, ROW).  Clip c of frame t is
    buffer[c // per, t, c % per]; `pose(t)` / `global_pos(t)` assemble one frame in clip order (a copy of that frame only)."""

    def __init__(self, buffer: torch.Tensor, n_clips: int):
        self.buffer, self.n_clips = buffer, n_clips
        self.world, self.n_frames, self.per = buffer.shape[0], buffer.shape[1], buffer.shape[2]

    def frame(self, t: int) -> torch.Tensor:
        return self.buffer[:, t].reshape(self.world * self.per, ROW)[: self.n_clips]

    def pose(self, t: int) -> torch.Tensor:
        return self.frame(t)[:, :88]

    def global_pos(self, t: int) -> torch.Tensor:
        return self.frame(t)[:, 88:91]

    def clip(self, c: int) -> torch.Tensor:
        """(T, ROW) rows of one clip: a view, no copy."""
        return self.buffer[c // self.per, :, c % self.per]


def gather_packed(rows: torch.Tensor, n_clips: int, dst: int = 0, group=None, out: torch.Tensor = None, async_op: bool = False):
    """Packed result rows of this rank's shard, (T, per, ROW) contiguous (rows beyond the shard's own clips are padding), to rank
    `dst` in ONE collective.  Returns (GatheredFrames on dst / None elsewhere, work handle or None).  With async_op the caller's
    stream is not blocked: call handle.wait() before reading the result (or before reusing `rows`)."""
    world, rank = _world_rank(group)
    assert rows.dim() == 3 and rows.shape[2] == ROW and rows.is_contiguous()
    T, per = rows.shape[0], rows.shape[1]
    assert per == shard_bounds(n_clips, world, rank)[2], "local shard size does not match shard_bounds"
    if world == 1:
        return GatheredFrames(rows.unsqueeze(0), n_clips), None
    if rank == dst:
        buf = out if out is not None else torch.empty((world, T, per, ROW), dtype=torch.float32, device=rows.device)
        assert buf.shape == (world, T, per, ROW) and buf.is_contiguous()
        work = dist.gather(rows, [buf[r] for r in range(world)], dst=dst, group=group, async_op=async_op)
        return GatheredFrames(buf, n_clips), work
    work = dist.gather(rows, None, dst=dst, group=group, async_op=async_op)
    return None, work


def gather_results(pose: torch.Tensor, gpos: torch.Tensor, n_clips: int, group=None, out: torch.Tensor = None):
    """One frame held as two arrays, to EVERY rank: all-gather of the local rows in rank order -> pose (n_clips,88),
    gpos (n_clips,3).  `out` (world*per, ROW) may be preallocated to avoid per-frame allocations."""
    world, rank = _world_rank(group)
    lo, hi, per = shard_bounds(n_clips, world, rank)
    assert pose.shape[0] == hi - lo, "local shard size does not match shard_bounds"
    rows = pack_rows(pose, gpos, per)
    if world == 1:
        full = rows
    else:
        full = out if out is not None else torch.empty((world * per, ROW), dtype=torch.float32, device=pose.device)
        dist.all_gather_into_tensor(full, rows, group=group)
    return full[:n_clips, :88], full[:n_clips, 88:91]


def gather_frames(pose: torch.Tensor, gpos: torch.Tensor, n_clips: int, dst: int = 0, group=None):
    """A whole batch of frames held as two arrays, local pose (T,n,88), gpos (T,n,3), to rank `dst` in one collective (SURVEY 8(e):
    "or once per run for all T frames").  Returns GatheredFrames on dst, None elsewhere.  Callers on the fast path let the engine
    write packed rows and use gather_packed directly; this wrapper pays one packing copy."""
    world, rank = _world_rank(group)
    lo, hi, per = shard_bounds(n_clips, world, rank)
    T, n = pose.shape[0], pose.shape[1]
    assert n == hi - lo, "local shard size does not match shard_bounds"
    rows = torch.zeros((T, per, ROW), dtype=torch.float32, device=pose.device)
    rows[:, :n, :88] = pose
    rows[:, :n, 88:91] = gpos
    got, _ = gather_packed(rows, n_clips, dst=dst, group=group)
    return got
