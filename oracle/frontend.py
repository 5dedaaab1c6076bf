"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement of the log-mel frontend the reference obtains from two librosa
calls (reference main.py:117-125; data/dataset.py:155-156,195-196):

    mel = librosa.feature.melspectrogram(y=chunk, sr=16000, n_mels=320, hop_length=512)
    mel = librosa.power_to_db(mel).astype(np.float32)

librosa (pinned only as ``librosa>=0.10.0``, reference requirements.txt:9) is a
third-party dependency that is absent from /root/reference and not installable
offline, so its published algorithm is restated here.

PARITY UNPINNED: the reference holds no tests or golden vectors for this stage
and librosa cannot be run here.  The restatement is cross-checked against
torchaudio's independent implementation (oracle/make_golden.py ->
tests/golden/frontend_torchaudio.npz; they differ only by fp32-vs-fp64 FFT) and
against ``transformers.audio_utils``, HuggingFace's numpy port of librosa's
filters.mel / melspectrogram / power_to_db (tests/test_oracle.py: filterbank
3.7e-9, dB spectrogram 7.6e-6 max).
"""
from __future__ import annotations

import numpy as np


def hz_to_mel_slaney(f):
    """Slaney mel scale (librosa ``htk=False``): linear below 1 kHz, log above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_t, mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr=16000, n_fft=2048, n_mels=320, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel defaults: slaney scale, slaney area norm, float32.
    Built in float64, returned float32 (n_mels, 1 + n_fft//2)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fftfreqs = np.arange(n_bins, dtype=np.float64) * (sr / n_fft)
    mel_pts = np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    W = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        W[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    W *= enorm[:, None]
    return W.astype(np.float32)


def hann_periodic(n_fft=2048) -> np.ndarray:
    """scipy.signal.get_window('hann', n_fft, fftbins=True), float64."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)


def power_spectrogram(y: np.ndarray, n_fft=2048, hop=512) -> np.ndarray:
    """|STFT|^2 with librosa>=0.10 defaults: center=True, pad_mode='constant',
    periodic Hann (float64) times float32 frames, float64 FFT rounded to
    complex64, power in float32.  Returns (1 + n_fft//2, T) float32."""
    y = np.asarray(y, dtype=np.float32)
    pad = n_fft // 2
    yp = np.concatenate([np.zeros(pad, np.float32), y, np.zeros(pad, np.float32)])
    T = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = yp[idx]                                   # (T, n_fft) float32
    w = hann_periodic(n_fft)
    X = np.fft.rfft(frames * w[None, :], axis=1).astype(np.complex64)   # (T, bins)
    P = (np.abs(X) ** 2).astype(np.float32)
    return np.ascontiguousarray(P.T)


def power_to_db(S: np.ndarray, amin=1e-10, top_db=80.0) -> np.ndarray:
    """librosa.power_to_db(ref=1.0): the floor is relative to the max of the
    whole array passed in one call, i.e. one chunk (reference main.py:125)."""
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, 1.0))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def logmel(y: np.ndarray, sr=16000, n_mels=320, hop=512, n_fft=2048, top_db=80.0,
           fb: np.ndarray | None = None) -> np.ndarray:
    """audio_to_mel's arithmetic (reference main.py:103-130) -> (n_mels, T) float32."""
    if fb is None:
        fb = mel_filterbank(sr, n_fft, n_mels)
    P = power_spectrogram(y, n_fft, hop)
    M = np.einsum("mf,ft->mt", fb, P, optimize=True).astype(np.float32)
    return power_to_db(M, top_db=top_db).astype(np.float32)


def logmel_f64(y: np.ndarray, sr=16000, n_mels=320, hop=512, n_fft=2048, top_db=80.0) -> np.ndarray:
    """All-float64 version (no complex64/float32 roundings) used to quantify
    how much of a mismatch is rounding noise of the recipe itself."""
    fb = mel_filterbank(sr, n_fft, n_mels).astype(np.float64)
    y = np.asarray(y, dtype=np.float64)
    pad = n_fft // 2
    yp = np.concatenate([np.zeros(pad), y, np.zeros(pad)])
    T = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    X = np.fft.rfft(yp[idx] * hann_periodic(n_fft)[None, :], axis=1)
    P = (X.real ** 2 + X.imag ** 2).T
    return power_to_db(fb @ P, top_db=top_db)
