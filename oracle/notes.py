"""ORACLE (test infrastructure, never shipped or measured as the product).

numpy restatement of the threshold -> piano-roll -> note grouping steps of the
reference (main.py:153-156 ``predict_chunk``; main.py:164-186
``combine_piano_rolls``; main.py:204-223 the per-pitch run grouping inside
``pianoroll_to_midi``; same code at scripts/evaluate.py:54-88).

Pinned by tests/golden/notes_reference.npz, produced by executing the
reference's own ``pianoroll_to_midi`` with a stand-in ``pretty_midi`` module
(oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def threshold_roll(probs: np.ndarray, threshold: float) -> np.ndarray:
    """(probs > threshold).float(): float32 compare against the threshold
    rounded to float32, strict."""
    p = np.asarray(probs, dtype=np.float32)
    return (p > np.float32(threshold)).astype(np.float32)


def combine_piano_rolls(rolls):
    """np.concatenate(axis=1); a single roll is returned as is (main.py:177-184)."""
    if len(rolls) == 1:
        return rolls[0]
    return np.concatenate(rolls, axis=1)


def group_notes(pianoroll: np.ndarray) -> np.ndarray:
    """Return int32 (n_notes, 3) rows (pitch_idx, onset_frame, offset_frame),
    pitch-major then onset-ascending -- the order main.py:204-223 appends
    notes in.  A note is a maximal run of ``> 0`` frames; offset is exclusive."""
    out = []
    for pitch_idx in range(pianoroll.shape[0]):
        active = pianoroll[pitch_idx] > 0
        changes = np.diff(np.concatenate([[0], active.astype(int), [0]]))
        onsets = np.where(changes == 1)[0]
        offsets = np.where(changes == -1)[0]
        for s, e in zip(onsets, offsets):
            if e / 1.0 > s / 1.0:          # `if end_time > start_time` (main.py:216), always true
                out.append((pitch_idx, int(s), int(e)))
    return np.asarray(out, dtype=np.int32).reshape(-1, 3)


def notes_to_events(notes: np.ndarray, fs: float, min_midi: int = 21):
    """(pitch, start_s, end_s, velocity) with times = idx / fs in float64
    (main.py:214-222; velocity fixed at 100)."""
    return [(int(min_midi + p), float(s) / fs, float(e) / fs, 100) for p, s, e in notes]
