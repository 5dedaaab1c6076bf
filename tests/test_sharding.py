"""CPU: chunk sharding + gather/stitch logic, incl. a real world_size-2 gloo run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from music_transcription_b200 import sharding
from oracle import notes as onotes


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 240, 241):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_range(240, 3, 8) == (90, 120)


def _rolls(n_chunks, T, seed):
    rng = np.random.default_rng(seed)
    rolls = (rng.random((n_chunks, 88, T)) < 0.15).astype(np.float32)
    rolls[:, 10, :] = 1.0          # a note sounding through every seam
    rolls[1, 20, T - 3:] = 1.0     # a note crossing the 1|2 seam only
    rolls[2, 20, :5] = 1.0
    return rolls


def test_stitch_equals_grouping_of_concatenated_roll():
    T, n = 50, 7
    rolls = _rolls(n, T, 0)
    want = onotes.group_notes(np.concatenate(list(rolls), axis=1))
    for world in (1, 2, 3, 7):
        parts = []
        for r in range(world):
            lo, hi = sharding.shard_range(n, r, world)
            local = onotes.group_notes(np.concatenate(list(rolls[lo:hi]), axis=1)) if hi > lo else np.zeros((0, 3), np.int32)
            local = local.copy()
            local[:, 1:] += lo * T
            parts.append(local)
        assert np.array_equal(sharding.stitch_notes(parts), want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rolls = _rolls(n, T, 1)
    lo, hi = sharding.shard_range(n, rank, world)
    local = onotes.group_notes(np.concatenate(list(rolls[lo:hi]), axis=1))
    got = sharding.gather_notes(local, lo * T)
    counts_local = np.arange((hi - lo) * 4 * 3, dtype=np.int64).reshape(hi - lo, 4, 3) + 1000 * rank
    allc = sharding.gather_counts(counts_local, n)
    q.put((rank, got, allc))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_notes_and_counts():
    T, n, world = 40, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = onotes.group_notes(np.concatenate(list(_rolls(n, T, 1)), axis=1))
    for rank, got, allc in res:
        assert np.array_equal(got, want)
        assert allc.shape == (n, 4, 3)
        lo, hi = sharding.shard_range(n, 1, world)
        assert allc[lo, 0, 0] == 1000 and allc[0, 0, 0] == 0
