"""Worker of tests/test_multigpu.py (launched by torch.distributed.run, one rank per GPU, NCCL): the N > 1 path of
DESIGN.md section 6 on real GPUs -- per-rank audio -> notes on a contiguous block of chunks, the note lists gathered
over NCCL and stitched, the threshold-sweep counts sharded by piece and all-gathered -- each compared bit-exactly with
what ONE rank computes for the whole input (and with the oracle)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from music_transcription_b200 import evaluate, pipeline, sharding, synth  # noqa: E402
from music_transcription_b200.transcription_model import TranscriptionModel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T = 938
    # ---- configs[3] shape in small: a recording of n chunks, rank r owns shard_range(n, r, world)
    n = 7
    sd = synth.synth_state_dict("cnn_rnn_large", 320, 128, 1, seed=4, gain=1.0)
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=128, num_layers=1, device=dev)
    m.load_state_dict(sd)
    wav_all = synth.cheap_wave_batch(n, 480000, seed=3)
    wav_all[2] *= 0.0                                              # a silent chunk: notes end / start at its seams
    lo, hi = sharding.shard_range(n, rank, world)
    wav = wav_all[lo:hi].to(dev)
    fe = pipeline.Frontend.get(device=dev)
    probs = torch.sigmoid(m(fe.logmel(wav)))
    cap = 88 * (((hi - lo) * T + 1) // 2)
    notes = torch.empty(cap, 3, dtype=torch.int32, device=dev)
    counts = torch.empty(89, dtype=torch.int32, device=dev)
    pipeline.extract_notes_async(probs, 0.5, notes, counts)
    got = sharding.gather_notes_device(notes, counts, lo * T)
    local_np = notes[:int(counts[88])].cpu().numpy()
    got_np = sharding.gather_notes(local_np, lo * T)               # the numpy-input variant (gloo-tested on CPU) over NCCL
    got_rolls = sharding.gather_rolls_notes(pipeline.pack_roll(probs, 0.5), T, n)   # the roll form: bits all-gathered, one grouping pass
    want, _ = pipeline.transcribe_chunks(m, wav_all.to(dev), threshold=0.5, batch=64)
    assert np.array_equal(got_rolls, want), (rank, len(got_rolls), len(want))
    # the streaming form: rolls handed over from host memory, exchange on a side stream, result collected later
    g = sharding.AsyncRollGather(hi - lo, n, T, dev)
    bits_host = pipeline.pack_roll(probs, 0.5).cpu()
    t0 = g.submit(bits_host)
    t1 = g.submit(bits_host.numpy())
    assert np.array_equal(g.result(t0), want) and np.array_equal(g.result(t1), want)
    g.ROWS_PER_CHUNK = 1                                           # force the "denser than the blind download" path
    g2 = sharding.AsyncRollGather(hi - lo, n, T, dev)
    g2.guess = 7
    for sl in g2.slots:
        sl["host_notes"] = sl["host_notes"][:7]
    assert np.array_equal(g2.result(g2.submit(bits_host)), want)
    assert len(want) > 50, len(want)
    assert np.array_equal(got, want), (rank, len(got), len(want))
    assert np.array_equal(got_np, want)
    # ---- configs[4]: 50 pieces x 100 thresholds, pieces sharded, int64 counts all-gathered: bit-exact
    n_pieces, thr = 50, np.linspace(0.01, 0.99, 100)
    lens = np.array([937, 938, 469] * 17, dtype=np.int32)[:n_pieces]
    P = [synth.planted_probs(88, T, thr, seed=i) for i in range(n_pieces)]
    Y = [synth.bernoulli_roll(88, T, 0.05, seed=i) for i in range(n_pieces)]
    plo, phi = sharding.shard_range(n_pieces, rank, world)
    Pl = torch.from_numpy(np.stack(P[plo:phi])).to(dev)
    Yl = torch.from_numpy(np.stack(Y[plo:phi])).to(dev)
    allc = sharding.gather_counts(evaluate.f1_counts_device(Pl, Yl, lens[plo:phi], thr), n_pieces)
    full = evaluate.f1_counts(torch.from_numpy(np.stack(P)).to(dev), torch.from_numpy(np.stack(Y)).to(dev), lens, thr)
    assert allc.dtype == np.int64 and allc.shape == (n_pieces, 100, 3) and np.array_equal(allc, full)
    if rank == 0:
        from oracle import f1 as of1
        for i in (0, 24, 25, 49):                                   # pieces on both sides of the rank seam
            assert np.array_equal(allc[i], of1.counts_grid([P[i]], [Y[i]], [lens[i]], thr)[0])
    dist.barrier()
    if rank == 0:
        print(f"NCCL_OK world={world} notes={len(want)} counts_sum={int(allc.sum())}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
