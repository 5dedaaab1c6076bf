"""Host-side mirror of the reference's evaluation kernel (scripts/evaluate.py).

    evaluate_at_threshold   scripts/evaluate.py:524-553
    run_threshold_tuning    scripts/evaluate.py:556-618   (same coarse-to-fine walk)

The reference re-runs the whole model for every threshold (37 forwards per
chunk by default).  Here the probabilities of a piece are computed once
(``probabilities``), the TP/FP/FN counts of every threshold of interest come from
one pass of ``amt_f1_counts`` on the GPU, and the walk is replayed on the host
from those integer counts -- bit-exact with sklearn's f1_score(zero_division=0).
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib


def f1_counts_device(probs: torch.Tensor, target: torch.Tensor, lengths, thresholds) -> torch.Tensor:
    """Launch-only form of ``f1_counts``: returns the int64 [n_pieces, n_thr, 3] table as a CUDA tensor without a host
    synchronisation (all launches on the current stream), so sweeps over many pieces / ranks pay one D2H at the end."""
    _lib.require_cuda(probs, "f1_counts probs")
    dev = probs.device
    probs = probs.float().contiguous()
    target = target.to(dev).float().contiguous()
    n_pieces, n_pitch, T = probs.shape
    thr64 = np.asarray(list(thresholds), dtype=np.float64).reshape(-1)
    thr32 = thr64.astype(np.float32)                       # torch compares in float32
    uniq, inverse = np.unique(thr32, return_inverse=True)  # sorted ascending
    lengths_t = lengths.to(dev, torch.int32) if torch.is_tensor(lengths) else torch.as_tensor(np.asarray(lengths, dtype=np.int32)).to(dev)
    out = torch.empty(n_pieces, len(uniq), 3, dtype=torch.int64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        for j0 in range(0, len(uniq), 512):
            chunk = torch.from_numpy(uniq[j0:j0 + 512].copy()).to(dev, non_blocking=True)
            one_block = len(uniq) <= 512
            for p0 in range(0, n_pieces, 65535):
                p1 = min(n_pieces, p0 + 65535)
                res = out[p0:p1] if one_block else torch.empty(p1 - p0, len(chunk), 3, dtype=torch.int64, device=dev)
                _lib.check(L.amt_f1_counts(_lib.ptr(probs[p0:p1]), _lib.ptr(target[p0:p1]), _lib.ptr(lengths_t[p0:p1]),
                                           p1 - p0, n_pitch, T, _lib.ptr(chunk), len(chunk), _lib.ptr(res),
                                           _lib.stream_ptr(dev)))
                if not one_block:
                    out[p0:p1, j0:j0 + len(chunk)] = res
    if len(inverse) == len(uniq) and np.array_equal(inverse, np.arange(len(uniq))):
        return out
    return out[:, torch.from_numpy(inverse.astype(np.int64)).to(dev)]


def f1_counts(probs: torch.Tensor, target: torch.Tensor, lengths, thresholds) -> np.ndarray:
    """probs/target (n_pieces, 88, T) CUDA float32, lengths (n_pieces,), thresholds: any order /
    duplicates allowed (float64 as np.arange yields them).  Returns int64 [n_pieces, n_thr, 3]
    = (TP, FP, FN) per piece and threshold, compare ``probs > float32(threshold)``.  One host sync (the final D2H)."""
    return f1_counts_device(probs, target, lengths, thresholds).cpu().numpy()


def f1_from_counts(counts: np.ndarray) -> np.ndarray:
    """sklearn binary F1 with zero_division=0 from (..., 3) integer counts, float64."""
    tp, fp, fn = counts[..., 0].astype(np.float64), counts[..., 1], counts[..., 2]
    den = 2 * tp + fp + fn
    return np.where(den == 0, 0.0, 2 * tp / np.where(den == 0, 1, den))


@torch.no_grad()
def probabilities(model, dataloader, device) -> tuple:
    """One forward per (mel, roll, lengths) batch item 0, as the reference scores it
    (scripts/evaluate.py:533-545).  Returns (probs [n, 88, Tmax], rolls, lengths)."""
    probs, rolls, lens = [], [], []
    L = _lib.lib()
    for mel, roll, lengths in dataloader:
        mel = mel.to(device)
        logits = model(mel)
        p = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.0, _lib.ptr(p), 0,
                                               _lib.stream_ptr(logits.device)))
        probs.append(p[0])
        rolls.append(roll[0].to(device).float())
        lens.append(int(lengths[0]))
    Tmax = max(p.shape[-1] for p in probs)
    P = torch.zeros(len(probs), 88, Tmax, device=device)
    Y = torch.zeros(len(probs), 88, Tmax, device=device)
    for i, (p, y) in enumerate(zip(probs, rolls)):
        P[i, :, :p.shape[-1]] = p
        Y[i, :, :y.shape[-1]] = y
    return P, Y, np.asarray(lens, dtype=np.int32)


@torch.no_grad()
def probabilities_bucketed(model, dataset, device, max_batch: int = 64) -> tuple:
    """``probabilities`` for an indexable dataset of (mel (1,n_mels,T), roll (88,T)) items, batched by exact
    length (cached.bucketed_batches): same (probs, rolls, lengths), in dataset order, as the one-at-a-time
    loop -- equal-length batches are bitwise batch-invariant -- with up to ``max_batch`` chunks per forward."""
    from . import cached
    n = len(dataset)
    per = [None] * n
    L = _lib.lib()
    for idx, mel, roll in cached.bucketed_batches(dataset, max_batch):
        mel = mel.to(device)
        logits = model(mel)
        p = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.0, _lib.ptr(p), 0,
                                               _lib.stream_ptr(logits.device)))
        roll = roll.to(device).float()
        for k, i in enumerate(idx):
            per[i] = (p[k], roll[k])
    Tmax = max(p.shape[-1] for p, _ in per) if n else 1
    P = torch.zeros(n, 88, Tmax, device=device)
    Y = torch.zeros(n, 88, Tmax, device=device)
    lens = np.zeros(n, dtype=np.int32)
    for i, (p, y) in enumerate(per):
        P[i, :, :p.shape[-1]] = p
        Y[i, :, :y.shape[-1]] = y
        lens[i] = p.shape[-1]
    return P, Y, lens


def evaluate_at_threshold(model, dataloader, dataset_info, threshold, _cache=None) -> float:
    """Mean framewise F1 over the loader at one threshold (reference signature)."""
    device = dataset_info["device"]
    P, Y, lens = _cache if _cache is not None else probabilities(model, dataloader, device)
    if len(lens) == 0:
        return 0.0
    c = f1_counts(P, Y, lens, [threshold])[:, 0]
    return float(np.mean(f1_from_counts(c)))


def threshold_schedule_walk(mean_f1_at, tune_range=(0.05, 0.95), tune_step=0.1, tune_min_step=0.01, tune_rounds=6):
    """The reference's coarse-to-fine schedule (scripts/evaluate.py:566-609): np.arange grid, strict
    ``>`` keeps the first best, window +-2*step clipped to [0.01, 0.99], step halves, stop below min_step."""
    tune_min, tune_max = tune_range
    step = tune_step
    best_threshold, best_f1 = 0.5, -1.0
    for _ in range(1, tune_rounds + 1):
        round_best_t, round_best_f1 = best_threshold, best_f1
        for t in np.arange(tune_min, tune_max + step / 2, step):
            f1 = mean_f1_at(t)
            if f1 > round_best_f1:
                round_best_f1, round_best_t = f1, t
        best_threshold, best_f1 = round_best_t, round_best_f1
        tune_min = max(0.01, best_threshold - 2 * step)
        tune_max = min(0.99, best_threshold + 2 * step)
        step = step / 2
        if step < tune_min_step:
            break
    return float(best_threshold), float(best_f1)


def run_threshold_tuning(args, model, dataloader, dataset_info):
    """Same result as the reference's run_threshold_tuning, one model pass instead of ~37."""
    device = dataset_info["device"]
    cache = probabilities(model, dataloader, device)
    P, Y, lens = cache
    memo = {}

    def mean_at(t):
        key = float(np.float32(t))
        if key not in memo:
            memo[key] = float(np.mean(f1_from_counts(f1_counts(P, Y, lens, [t])[:, 0]))) if len(lens) else 0.0
        return memo[key]

    return threshold_schedule_walk(mean_at, tuple(args.tune_range), args.tune_step, args.tune_min_step, args.tune_rounds)
