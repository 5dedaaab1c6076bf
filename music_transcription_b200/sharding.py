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

import contextlib
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


def _merge_at(notes: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Rows ``idx`` (sorted) continue the row before them: drop them and hand their offset to the surviving row."""
    if len(idx) == 0:
        return notes
    run_end = np.append(np.diff(idx) != 1, True)              # last row of each run of consecutive merged rows
    run_start = np.append(True, run_end[:-1])
    out = notes.copy()
    out[idx[run_start] - 1, 2] = notes[idx[run_end], 2]
    return np.delete(out, idx, axis=0)


def merge_touching(notes: np.ndarray) -> np.ndarray:
    """notes int (n,3) sorted by (pitch, onset): merge runs of consecutive rows of one pitch whose offset == the next
    onset (a note cut by one or more shard seams).  Vectorised: a 2-hour recording has ~10^5 notes."""
    notes = np.asarray(notes).reshape(-1, 3)
    if len(notes) < 2:
        return notes
    cont = (notes[1:, 0] == notes[:-1, 0]) & (notes[1:, 1] == notes[:-1, 2])        # row i+1 continues row i
    return _merge_at(notes, np.flatnonzero(cont) + 1)


def stitch_notes(per_rank, per_pitch_counts=None) -> np.ndarray:
    """per_rank: list (rank / batch order = time order) of int32 (n_r,3) note arrays with GLOBAL frame indices, each
    pitch-major / onset-ascending.  Returns the note list of the concatenated roll: the parts are interleaved pitch by
    pitch (slices located by the per-pitch note counts -- ``per_pitch_counts`` as amt_threshold_notes reports them, else
    recounted here) and only the rows at part boundaries, the one place a cut note can sit, are tested for merging."""
    per_rank = [np.asarray(a, dtype=np.int32).reshape(-1, 3) for a in per_rank]
    if not per_rank or sum(len(a) for a in per_rank) == 0:
        return np.zeros((0, 3), np.int32)
    if len(per_rank) == 1:
        return merge_touching(per_rank[0])
    if per_pitch_counts is None:
        n_pitch = max(int(a[:, 0].max()) + 1 for a in per_rank if len(a))
        per_pitch_counts = [np.bincount(a[:, 0], minlength=n_pitch) for a in per_rank]
    c = np.stack([np.asarray(x, dtype=np.int64) for x in per_pitch_counts])        # [parts][pitches]
    offs = np.concatenate([np.zeros((len(c), 1), np.int64), np.cumsum(c, axis=1)], axis=1)
    n_pitch = c.shape[1]
    allnotes = np.concatenate([a[o[p]:o[p + 1]] for p in range(n_pitch) for a, o in zip(per_rank, offs)], axis=0)
    seg = c.T.reshape(-1)                                                           # segment sizes in (pitch, part) order
    start = np.cumsum(seg) - seg
    part = np.tile(np.arange(len(c)), n_pitch)
    cand = start[(seg > 0) & (part > 0) & (start > 0)]                              # first row of every later part's segment
    cand = cand[(allnotes[cand, 0] == allnotes[cand - 1, 0]) & (allnotes[cand, 1] == allnotes[cand - 1, 2])]
    return _merge_at(allnotes, cand)


def _device_for_backend():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


_SIDE = {}


def _side_stream(dev):
    """Collectives whose inputs come from the HOST have no business waiting behind whatever compute the caller has
    already queued on its stream (NCCL orders itself after the *current* stream): they run under a side stream."""
    if dev.type != "cuda":
        return None
    key = str(dev)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev)
    return _SIDE[key]


def gather_notes(local_notes: np.ndarray, frame_offset: int) -> np.ndarray:
    """All ranks call this with their local note list (host array, frame indices local to their block) and the
    global frame index of their first frame.  Every rank returns the stitched global list."""
    local = np.asarray(local_notes, dtype=np.int32).reshape(-1, 3).copy()
    local[:, 1:] += np.int32(frame_offset)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_touching(local)
    dev = _device_for_backend()
    world = dist.get_world_size()
    n_pitch = 88 if not len(local) else max(88, int(local[:, 0].max()) + 1)
    side = _side_stream(dev)
    ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
    with ctx:
        # per-pitch counts travel with the lists: the merge on the host is then a slice interleave, not a sort
        head = torch.zeros(1 + 1024, dtype=torch.int64)
        head[0] = len(local)
        head[1:1 + n_pitch] = torch.from_numpy(np.bincount(local[:, 0], minlength=n_pitch).astype(np.int64))
        head = head.to(dev)
        heads = [torch.zeros_like(head) for _ in range(world)]
        dist.all_gather(heads, head)
        heads = torch.stack(heads).cpu().numpy()
        counts = [int(h[0]) for h in heads]
        cap = max(max(counts), 1)
        buf = torch.zeros(cap, 3, dtype=torch.int32)
        if len(local):
            buf[:len(local)] = torch.from_numpy(local)
        buf = buf.to(dev)
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        parts = [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]
    n_pitch = int(max(np.flatnonzero(heads[:, 1:].sum(0) > 0).max(initial=0) + 1, 1))
    return stitch_notes(parts, [h[1:1 + n_pitch] for h in heads])


def gather_notes_device(notes_dev: torch.Tensor, counts_dev: torch.Tensor, frame_offset: int) -> np.ndarray:
    """``gather_notes`` for the output of ``amt_threshold_notes`` as it lies on the GPU: ``notes_dev`` int32 (cap,3) and
    ``counts_dev`` int32 (n_pitch+1: per-pitch counts, then the total).  The per-pitch counts travel with the lists, so
    the host merge is a pitch-wise slice concatenation (no sort).  Two small collectives (counts, then padded lists)
    over NCCL / NVLink; every rank returns the stitched global list."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    n_pitch = counts_dev.numel() - 1
    if world == 1:
        counts = counts_dev.cpu().numpy()
        local = notes_dev[:int(counts[n_pitch])].cpu().numpy().copy()
        local[:, 1:] += np.int32(frame_offset)
        return merge_touching(local)
    allc = torch.empty(world, n_pitch + 1, dtype=torch.int32, device=counts_dev.device)
    dist.all_gather_into_tensor(allc, counts_dev.contiguous())
    allc = allc.cpu().numpy()
    totals = allc[:, n_pitch]
    cap = max(int(totals.max()), 1)
    if cap > notes_dev.shape[0]:
        raise RuntimeError(f"gather_notes_device: a rank holds {cap} notes, more than the {notes_dev.shape[0]}-row buffers")
    send = notes_dev[:cap].clone()                       # own rows beyond the local total are never read back
    send[:, 1:] += int(frame_offset)
    allb = torch.empty(world, cap, 3, dtype=torch.int32, device=notes_dev.device)
    dist.all_gather_into_tensor(allb, send)
    allb = allb.cpu().numpy()
    return stitch_notes([allb[r, :totals[r]] for r in range(world)], [allc[r, :n_pitch] for r in range(world)])


def gather_rolls_notes(bits_local: torch.Tensor, T: int, n_total: int) -> np.ndarray:
    """The N > 1 exchange of DESIGN.md section 6 in its roll form: ``bits_local`` int32
    (n_local, 88, ceil(T/32)) CUDA = this rank's bit-packed rolls (its ``shard_range`` of the ``n_total`` chunks).
    ONE all-gather of 10.6 KB per chunk over NCCL / NVLink, then one grouping pass over the gathered roll on the GPU
    (``amt_bits_notes``): the note list of the whole recording, seams included, with no host-side merging.  Every rank
    returns it.  Single process: just the grouping pass."""
    from . import pipeline
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return pipeline.extract_notes_from_bits(bits_local, T)
    n_pitch, words = bits_local.shape[1], bits_local.shape[2]
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    cap = max(sizes)
    send = bits_local.contiguous()
    if send.shape[0] < cap:                                                       # uneven shards: pad the short ones
        send = torch.cat([send, torch.zeros(cap - send.shape[0], n_pitch, words, dtype=send.dtype, device=send.device)])
    allb = torch.empty(world, cap, n_pitch, words, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(allb, send)
    rolls = allb.view(world * cap, n_pitch, words) if min(sizes) == cap else torch.cat([allb[r, :sizes[r]] for r in range(world)])
    return pipeline.extract_notes_from_bits(rolls, T)


class AsyncRollGather:
    """``gather_rolls_notes`` for a STREAMING host: the packed rolls of a finished recording are handed over from pinned
    host memory (where the streaming path has just delivered them), and upload, all-gather, grouping pass and download of
    the note list all run on a side stream while the caller's stream is already busy with the next recording.  The host
    never waits for work queued behind future batches: ``submit`` returns at once, ``result`` is asked for one recording
    later (two slots).

        g = AsyncRollGather(n_local, n_total, T, device)
        t = g.submit(bits_host)        # int32 (n_local, 88, ceil(T/32)) host tensor / array
        ...                            # launch the next recording's batches
        notes = g.result(t)            # int32 (n, 3) numpy: the whole recording's note list
    """

    ROWS_PER_CHUNK = 1024              # rows of the note list downloaded blindly; more -> one extra blocking copy

    def __init__(self, n_local: int, n_total: int, T: int, device, n_pitch: int = 88):
        self.n_local, self.n_total, self.T, self.n_pitch = n_local, n_total, T, n_pitch
        self.dev = torch.device(device)
        self.words = (T + 31) // 32
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.sizes = [shard_range(n_total, r, self.world)[1] - shard_range(n_total, r, self.world)[0] for r in range(self.world)]
        if self.sizes[self.rank] != n_local:
            raise ValueError("AsyncRollGather: n_local is not this rank's shard_range of n_total")
        self.side = torch.cuda.Stream(self.dev)
        cap_chunks = max(self.sizes)
        self.cap = n_pitch * ((n_total * T + 1) // 2)
        self.guess = min(self.cap, self.ROWS_PER_CHUNK * n_total)
        self.slots = []
        for _ in range(2):
            sl = {"host_bits": torch.zeros(cap_chunks, n_pitch, self.words, dtype=torch.int32).pin_memory(),
                  "bits": torch.zeros(cap_chunks, n_pitch, self.words, dtype=torch.int32, device=self.dev),
                  "all": torch.empty(self.world, cap_chunks, n_pitch, self.words, dtype=torch.int32, device=self.dev),
                  "notes": torch.empty(self.cap, 3, dtype=torch.int32, device=self.dev),
                  "counts": torch.empty(n_pitch + 1, dtype=torch.int32, device=self.dev),
                  "scratch": torch.empty(2 * n_pitch * n_total, dtype=torch.int32, device=self.dev),
                  "host_notes": torch.empty(self.guess, 3, dtype=torch.int32).pin_memory(),
                  "host_counts": torch.zeros(n_pitch + 1, dtype=torch.int32).pin_memory(),
                  "done": torch.cuda.Event(), "busy": False}
            self.slots.append(sl)
        self.n = 0

    def submit(self, bits, after=()) -> int:
        """``bits``: this rank's packed rolls int32 (n_local, 88, words) -- a HOST tensor / array (copied into a pinned
        staging buffer and uploaded), or a CUDA tensor (used in place: keep it unchanged until ``result``).  ``after``:
        CUDA events the exchange must wait for (e.g. the completion of the lanes that wrote a CUDA ``bits``); for a CUDA
        ``bits`` the caller's current stream is always waited for."""
        from . import _lib
        ticket = self.n
        sl = self.slots[ticket & 1]
        if sl["busy"]:
            raise RuntimeError("AsyncRollGather: collect result(ticket - 2) before submitting again")
        on_device = torch.is_tensor(bits) and bits.is_cuda
        if on_device:
            local = bits.view(torch.int32).reshape(self.n_local, self.n_pitch, self.words)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.dev))
            after = tuple(after) + (ready,)
        else:
            src = bits if torch.is_tensor(bits) else torch.from_numpy(np.ascontiguousarray(bits))
            sl["host_bits"][:self.n_local].copy_(src.view(torch.int32).reshape(self.n_local, self.n_pitch, self.words))
        with torch.cuda.device(self.dev), torch.cuda.stream(self.side):
            for ev in after:
                self.side.wait_event(ev)
            if on_device:
                cap = sl["bits"].shape[0]
                send = local if (self.world == 1 or self.n_local == cap) else None
                if send is None:
                    sl["bits"][:self.n_local].copy_(local, non_blocking=True)
                    send = sl["bits"]
            else:
                sl["bits"].copy_(sl["host_bits"], non_blocking=True)
                send = sl["bits"]
            if self.world > 1:
                dist.all_gather_into_tensor(sl["all"], send.contiguous())
                cap = sl["all"].shape[1]
                rolls = (sl["all"].view(self.world * cap, self.n_pitch, self.words) if min(self.sizes) == cap
                         else torch.cat([sl["all"][r, :self.sizes[r]] for r in range(self.world)]))
            else:
                rolls = send[:self.n_local]
            _lib.check(_lib.lib().amt_bits_notes(_lib.ptr(rolls), self.n_total, self.n_pitch, self.T, _lib.ptr(sl["notes"]), self.cap,
                                                 _lib.ptr(sl["counts"]), _lib.ptr(sl["scratch"]), sl["scratch"].numel(),
                                                 self.side.cuda_stream))
            sl["host_counts"].copy_(sl["counts"], non_blocking=True)
            sl["host_notes"].copy_(sl["notes"][:self.guess], non_blocking=True)
            sl["done"].record(self.side)
            sl["keepalive"] = (rolls, send)
        sl["busy"] = True
        self.n += 1
        return ticket

    def result(self, ticket: int) -> np.ndarray:
        sl = self.slots[ticket & 1]
        if not sl["busy"]:
            raise RuntimeError("AsyncRollGather: no submission pending in this slot")
        sl["done"].synchronize()
        total = int(sl["host_counts"][self.n_pitch])
        if total <= self.guess:
            out = sl["host_notes"][:total].numpy().copy()
        else:                                                      # denser than ROWS_PER_CHUNK notes per chunk: fetch the rest
            with torch.cuda.device(self.dev), torch.cuda.stream(self.side):
                out = sl["notes"][:total].cpu().numpy()
        sl["busy"] = False
        return out


def gather_counts(local_counts, n_total: int) -> np.ndarray:
    """local_counts int64 [n_local, n_thr, 3] (numpy, or a tensor on the backend's device) for this rank's
    ``shard_range`` of ``n_total`` pieces -> int64 [n_total, n_thr, 3] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_counts.cpu().numpy() if torch.is_tensor(local_counts) else np.asarray(local_counts, dtype=np.int64)
    dev = _device_for_backend()
    world = dist.get_world_size()
    lc = local_counts if torch.is_tensor(local_counts) else torch.from_numpy(np.asarray(local_counts, dtype=np.int64))
    n_thr = lc.shape[1]
    cap = -(-n_total // world)
    buf = torch.zeros(cap, n_thr, 3, dtype=torch.int64, device=dev)
    buf[:len(lc)] = lc.to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    parts = []
    for r, b in enumerate(bufs):
        lo, hi = shard_range(n_total, r, world)
        parts.append(b[:hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)


def batch_ranges(n: int, max_batch: int):
    """Split ``n`` chunks into the fewest batches of at most ``max_batch``, sizes balanced (240 @ 64 -> 4 x 60)."""
    if n <= 0:
        return []
    k = -(-n // max_batch)
    return [shard_range(n, i, k) for i in range(k)]
