"""Chunk-level data parallelism over the GPUs of one box (one process per GPU).

The reference has no distributed code; what shards is the unit
main.py already produces: independent 30-s chunks (fresh LSTM state, per-chunk attention
and top_db floor, main.py:258-266).  Rank r owns a contiguous block of chunks, runs the whole
audio->roll->notes path locally with no data-path collective, and the only exchange is one
small gather at the end (NCCL over NVLink on GPUs, gloo in the CPU tests):

  * ``gather_notes``: per-rank note lists (frame indices already global) are all-gathered and
    stitched; a note that crosses a rank seam was emitted as two touching notes
    (left.offset == right.onset), which only happens at seams, so they are merged -- the
    result equals grouping the concatenated roll as main.py:270-275 does.
  * ``gather_counts``: per-piece TP/FP/FN tables of the threshold sweep (pieces sharded).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` units for ``rank`` (sizes differ by at most 1), so only
    world-1 seams cross ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_touching(notes: np.ndarray) -> np.ndarray:
    """notes int (n,3) sorted by (pitch, onset): merge consecutive rows of one pitch whose
    offset == next onset (a note cut by a shard seam)."""
    if len(notes) == 0:
        return notes.reshape(0, 3)
    out = [list(notes[0])]
    for p, s, e in notes[1:]:
        last = out[-1]
        if p == last[0] and s == last[2]:
            last[2] = e
        else:
            out.append([p, s, e])
    return np.asarray(out, dtype=notes.dtype).reshape(-1, 3)


def stitch_notes(per_rank) -> np.ndarray:
    """per_rank: list (rank order) of int32 (n_r,3) note arrays with GLOBAL frame indices, each
    pitch-major / onset-ascending.  Returns the note list of the concatenated roll."""
    per_rank = [np.asarray(a, dtype=np.int32).reshape(-1, 3) for a in per_rank]
    allnotes = np.concatenate(per_rank, axis=0) if per_rank else np.zeros((0, 3), np.int32)
    if len(allnotes) == 0:
        return allnotes
    order = np.lexsort((allnotes[:, 1], allnotes[:, 0]))        # by pitch, then onset (ranks are time-ordered)
    return merge_touching(allnotes[order])


def _device_for_backend():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_notes(local_notes: np.ndarray, frame_offset: int) -> np.ndarray:
    """All ranks call this with their local note list (frame indices local to their block) and the
    global frame index of their first frame.  Every rank returns the stitched global list."""
    local = np.asarray(local_notes, dtype=np.int32).reshape(-1, 3).copy()
    local[:, 1:] += np.int32(frame_offset)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_touching(local)
    dev = _device_for_backend()
    world = dist.get_world_size()
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros(cap, 3, dtype=torch.int32, device=dev)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    return stitch_notes([b[:c].cpu().numpy() for b, c in zip(bufs, counts)])


def gather_counts(local_counts: np.ndarray, n_total: int) -> np.ndarray:
    """local_counts int64 [n_local, n_thr, 3] for this rank's ``shard_range`` of ``n_total`` pieces ->
    int64 [n_total, n_thr, 3] on every rank."""
    local_counts = np.asarray(local_counts, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_counts
    dev = _device_for_backend()
    world, rank = dist.get_world_size(), dist.get_rank()
    n_thr = local_counts.shape[1]
    cap = -(-n_total // world)
    buf = torch.zeros(cap, n_thr, 3, dtype=torch.int64, device=dev)
    buf[:len(local_counts)] = torch.from_numpy(local_counts).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    parts = []
    for r, b in enumerate(bufs):
        lo, hi = shard_range(n_total, r, world)
        parts.append(b[:hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)
