"""Cached-chunk datasets in front of the evaluation path.

The reference pre-processes MAESTRO into ``<cache_dir>/<split>/chunk_%06d.pt`` files holding
``{'mel': (1, n_mels, T) f32, 'roll': (88, T) f32}`` plus ``<cache_dir>/<split>_metadata.pkl``
(scripts/preprocess_dataset.py:66-69, :138-154) and reads them back with ``CachedMaestroDataset``
(data/cached_dataset.py:11-88); ``scripts/evaluate.py`` then scores one item at a time
(``DataLoader(batch_size=1, collate_fn=collate_fn)``, :313-318 and train/train_transcriber.py:23-39).

This module reads and writes the same on-disk format and adds the batching the reference lacks:
**equal-length bucketing**.  A chunk may only share a batch with chunks of exactly the same frame count --
zero-padding a shorter chunk to a common length would change its result (the padded frames carry
``relu(bias)`` after the first folded BatchNorm and feed the backward LSTM), whereas batches of equal length
are bitwise batch-invariant on this library (tests/test_gpu_model.py).  ``bucketed_batches`` therefore yields
``(indices, mel (B,1,n_mels,T), roll (B,88,T))`` per exact length, and ``evaluate.probabilities_bucketed``
returns exactly what the one-at-a-time loop returns, in dataset order.
"""
from __future__ import annotations

import os
import pickle
from collections import OrderedDict
from typing import Iterable, List, Sequence, Tuple

import torch


class CachedChunkDataset:
    """Mel-format ``CachedMaestroDataset`` (data/cached_dataset.py:11-88): item i -> (mel (1,n_mels,T), roll (88,T))."""

    def __init__(self, cache_dir: str = "cached_dataset", split: str = "train"):
        self.cache_dir, self.split = cache_dir, split
        self.split_cache_dir = os.path.join(cache_dir, split)
        metadata_path = os.path.join(cache_dir, f"{split}_metadata.pkl")
        if not os.path.exists(metadata_path):
            raise FileNotFoundError(f"Cache not found at {metadata_path}. Run preprocess_dataset.py first!")
        with open(metadata_path, "rb") as f:
            self.metadata = pickle.load(f)
        self.num_chunks = self.metadata["num_chunks"]
        if not os.path.exists(self.split_cache_dir):
            raise FileNotFoundError(f"Cache directory not found: {self.split_cache_dir}. Run preprocess_dataset.py first!")

    def __len__(self):
        return self.num_chunks

    def path(self, idx: int) -> str:
        return os.path.join(self.split_cache_dir, f"chunk_{idx:06d}.pt")

    def __getitem__(self, idx: int):
        p = self.path(idx)
        if not os.path.exists(p):
            raise FileNotFoundError(f"Cached chunk not found: {p}. Re-run preprocess_dataset.py")
        data = torch.load(p, weights_only=False)
        if "mel" not in data:
            raise ValueError(f"{p}: waveform / token caches belong to the AST model, which is out of scope here")
        return data["mel"], data["roll"]

    def frames(self, idx: int) -> int:
        return int(self[idx][0].shape[-1])


def write_cache(cache_dir: str, split: str, items: Sequence[Tuple[torch.Tensor, torch.Tensor]], chunk_length=30.0,
                overlap=0.0, sr=16000, n_mels=320, hop_length=512) -> None:
    """Write ``items`` = [(mel (1,n_mels,T), roll (88,T)), ...] in the reference's cache layout."""
    d = os.path.join(cache_dir, split)
    os.makedirs(d, exist_ok=True)
    for i, (mel, roll) in enumerate(items):
        torch.save({"mel": mel.float().cpu(), "roll": roll.float().cpu()}, os.path.join(d, f"chunk_{i:06d}.pt"))
    meta = {"root_dir": None, "chunk_length": chunk_length, "overlap": overlap, "split": split, "num_chunks": len(items),
            "chunks": None, "sr": sr, "n_mels": n_mels, "hop_length": hop_length, "return_waveform": False,
            "tokenize": False, "data_type": "mel"}
    with open(os.path.join(cache_dir, f"{split}_metadata.pkl"), "wb") as f:
        pickle.dump(meta, f)


def collate_fn(batch):
    """train/train_transcriber.py:23-39: pad to the longest item, return (mel, roll, lengths)."""
    mels, rolls = zip(*batch)
    lengths = [m.shape[-1] for m in mels]
    max_T = max(lengths)
    mel = torch.stack([torch.nn.functional.pad(m, (0, max_T - m.shape[-1])) for m in mels])
    roll = torch.stack([torch.nn.functional.pad(r, (0, max_T - r.shape[-1])) for r in rolls])
    return mel, roll, torch.tensor(lengths, dtype=torch.long)


def bucketed_batches(dataset, max_batch: int = 64, indices: Iterable[int] | None = None):
    """Yield (indices, mel (B,1,n_mels,T), roll (B,88,T)) with every batch holding chunks of ONE exact length,
    buckets in order of first appearance, at most ``max_batch`` chunks each."""
    buckets: "OrderedDict[int, List[int]]" = OrderedDict()
    cache = {}
    for i in (range(len(dataset)) if indices is None else indices):
        mel, roll = dataset[i]
        cache[i] = (mel, roll)
        buckets.setdefault(int(mel.shape[-1]), []).append(i)
    for T, idx in buckets.items():
        for j in range(0, len(idx), max_batch):
            part = idx[j:j + max_batch]
            mel = torch.stack([cache[i][0] for i in part])
            roll = torch.stack([cache[i][1] for i in part])
            yield part, mel, roll
