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


def group_notes_onset_aware(frame: np.ndarray, onset: np.ndarray, offset: np.ndarray | None = None) -> np.ndarray:
    """Onset / offset-aware decoding of the three thresholded heads of CNNRNNModelLarge (reference
    models/cnn_rnn_model.py:333-345 computes them; the reference's inference path uses the frame head only and has no
    decoder for the other two -- SURVEY.md section 8f rank 4).  The RULE is this repository's (amt.h, amt_onset_notes);
    this loop is its definition and the GPU kernel must reproduce it bit for bit.  Inputs (n_pitch, n_frames) arrays,
    non-zero = active.  Per pitch: a rising edge of the onset roll starts a note at t; the note ends at the first t' > t
    where neither frame nor onset is active, or offset is active, or another onset rises -- or at n_frames.  Frames that
    no onset opened are ignored.  Returns int32 (n, 3) rows (pitch, onset, offset-exclusive), pitch-major, onset ascending."""
    F = np.asarray(frame) > 0
    ON = np.asarray(onset) > 0
    OFF = np.zeros_like(F) if offset is None else np.asarray(offset) > 0
    out = []
    n_frames = F.shape[1]
    for p in range(F.shape[0]):
        start = -1
        for t in range(n_frames):
            rising = ON[p, t] and not (t > 0 and ON[p, t - 1])
            boundary = (not (F[p, t] or ON[p, t])) or rising or OFF[p, t]
            if start >= 0 and boundary:
                out.append((p, start, t))
                start = -1
            if rising:
                start = t
        if start >= 0:
            out.append((p, start, n_frames))
    return np.asarray(out, dtype=np.int32).reshape(-1, 3)
