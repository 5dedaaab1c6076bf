"""ORACLE (test infrastructure, never shipped or measured as the product).

numpy restatement of the framewise-F1 arithmetic of the reference's evaluation
(scripts/evaluate.py:524-553 ``evaluate_at_threshold``; :556-618
``run_threshold_tuning``): per chunk, threshold the probabilities with a strict
float32 ``>``, crop to the valid length, flatten, and score with
``sklearn.metrics.f1_score(zero_division=0)`` = 2TP / (2TP + FP + FN), 0 when
the denominator is 0; the sweep keeps the mean over chunks.

Pinned by tests/golden/f1_reference.npz, produced by executing the reference's
own two functions with a stub model (oracle/make_golden.py), and checked
against sklearn directly in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np


def counts(probs: np.ndarray, target: np.ndarray, length: int, threshold: float):
    """(TP, FP, FN) for one chunk: probs/target (88, T), first ``length`` frames."""
    p = np.asarray(probs, dtype=np.float32)[:, :length] > np.float32(threshold)
    y = np.asarray(target)[:, :length] > 0.5
    tp = int(np.count_nonzero(p & y))
    fp = int(np.count_nonzero(p & ~y))
    fn = int(np.count_nonzero(~p & y))
    return tp, fp, fn


def counts_grid(probs_list, target_list, lengths, thresholds) -> np.ndarray:
    """int64 [n_pieces, n_thr, 3]."""
    out = np.zeros((len(probs_list), len(thresholds), 3), dtype=np.int64)
    for i, (p, y, L) in enumerate(zip(probs_list, target_list, lengths)):
        for j, t in enumerate(thresholds):
            out[i, j] = counts(p, y, int(L), float(t))
    return out


def f1_from_counts(tp, fp, fn) -> float:
    """sklearn binary f1 with zero_division=0."""
    den = 2 * int(tp) + int(fp) + int(fn)
    return 0.0 if den == 0 else (2.0 * int(tp)) / den


def mean_f1(counts_ij: np.ndarray) -> float:
    """Unweighted mean of per-chunk F1 (scripts/evaluate.py:553)."""
    if len(counts_ij) == 0:
        return 0.0
    return float(np.mean([f1_from_counts(*c) for c in counts_ij]))


def threshold_walk(mean_f1_at, tune_range=(0.05, 0.95), step=0.1, min_step=0.01, rounds=6):
    """The coarse-to-fine walk of run_threshold_tuning (scripts/evaluate.py:566-609).
    ``mean_f1_at(t)`` plays evaluate_at_threshold.  Returns (best_t, best_f1, visited)."""
    tune_min, tune_max = tune_range
    best_threshold, best_f1 = 0.5, -1.0
    visited = []
    for _ in range(1, rounds + 1):
        round_best_t, round_best_f1 = best_threshold, best_f1
        for t in np.arange(tune_min, tune_max + step / 2, step):
            f1 = mean_f1_at(t)
            visited.append(float(t))
            if f1 > round_best_f1:
                round_best_f1, round_best_t = f1, t
        best_threshold, best_f1 = round_best_t, round_best_f1
        tune_min = max(0.01, best_threshold - 2 * step)
        tune_max = min(0.99, best_threshold + 2 * step)
        step = step / 2
        if step < min_step:
            break
    return float(best_threshold), float(best_f1), visited
