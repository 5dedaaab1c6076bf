"""Cached-chunk format and equal-length bucketing."""
import pickle

import numpy as np
import pytest
import torch

from music_transcription_b200 import cached


def _items(lengths, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(1, 320, T, generator=g) * 20 - 30, (torch.rand(88, T, generator=g) < 0.05).float()) for T in lengths]


def test_cache_round_trip_in_the_reference_layout(tmp_path):
    items = _items([938, 938, 469, 200])
    cached.write_cache(str(tmp_path), "test", items)
    meta = pickle.load(open(tmp_path / "test_metadata.pkl", "rb"))
    assert meta["num_chunks"] == 4 and meta["chunk_length"] == 30.0 and meta["data_type"] == "mel"
    raw = torch.load(tmp_path / "test" / "chunk_000002.pt", weights_only=False)
    assert set(raw) == {"mel", "roll"} and raw["mel"].shape == (1, 320, 469) and raw["roll"].shape == (88, 469)
    ds = cached.CachedChunkDataset(str(tmp_path), "test")
    assert len(ds) == 4
    for i, (mel, roll) in enumerate(items):
        m, r = ds[i]
        assert torch.equal(m, mel) and torch.equal(r, roll)
    with pytest.raises(FileNotFoundError):
        cached.CachedChunkDataset(str(tmp_path), "validation")


def test_bucketing_groups_exact_lengths_only_and_covers_every_index():
    lengths = [938, 469, 938, 200, 938, 469, 938, 938]
    items = _items(lengths, seed=1)
    seen = []
    for idx, mel, roll in cached.bucketed_batches(items, max_batch=3):
        Ts = {lengths[i] for i in idx}
        assert len(Ts) == 1 and mel.shape == (len(idx), 1, 320, Ts.pop()) and roll.shape[0] == len(idx) and len(idx) <= 3
        for k, i in enumerate(idx):
            assert torch.equal(mel[k], items[i][0])
        seen += idx
    assert sorted(seen) == list(range(len(lengths)))
    mel, roll, lens = cached.collate_fn([items[0], items[1]])
    assert mel.shape == (2, 1, 320, 938) and lens.tolist() == [938, 469] and float(mel[1, ..., 469:].abs().max()) == 0.0


@pytest.mark.gpu
def test_bucketed_probabilities_equal_the_one_at_a_time_loop(tmp_path):
    from torch.utils.data import DataLoader
    from music_transcription_b200 import evaluate, synth
    from music_transcription_b200.transcription_model import TranscriptionModel
    lengths = [938, 469, 938, 200, 938, 469, 938]
    cached.write_cache(str(tmp_path), "test", _items(lengths, seed=2))
    ds = cached.CachedChunkDataset(str(tmp_path), "test")
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=128, num_layers=1, device="cuda")
    m.load_state_dict(synth.synth_state_dict("cnn_rnn_large", 320, 128, 1, seed=3))
    loader = DataLoader(ds, batch_size=1, shuffle=False, collate_fn=cached.collate_fn)
    P1, Y1, l1 = evaluate.probabilities(m, loader, "cuda")
    P2, Y2, l2 = evaluate.probabilities_bucketed(m, ds, "cuda", max_batch=4)
    assert np.array_equal(l1, l2) and torch.equal(Y1, Y2) and torch.equal(P1, P2)
    c1 = evaluate.f1_counts(P1, Y1, l1, np.linspace(0.05, 0.95, 19))
    c2 = evaluate.f1_counts(P2, Y2, l2, np.linspace(0.05, 0.95, 19))
    assert np.array_equal(c1, c2)
