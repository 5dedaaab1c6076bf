"""Deterministic synthetic inputs for the audio->piano-roll path.

No dataset and no trained checkpoint exist offline, so every test, the golden
generator (oracle/make_golden.py) and bench.py build their inputs here:

* ``piano_chord``      -- the SURVEY.md section 8(d) synthetic 30-s "piano chord".
* ``state_dict_spec``  -- the exact key/shape list of a reference checkpoint
                          (reference
                          models/cnn_rnn_model.py:28-55,179-260).
* ``synth_state_dict`` -- seeded weights for that key list with non-trivial
                          BatchNorm statistics, so BN folding is really tested.
* ``planted_probs``    -- probability rolls with values planted exactly on
                          float32(threshold) for the strict ``>`` compare.

Everything is generated with CPU generators, so the same arrays appear in
this container and on the GPU box.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch

SR = 16000
CHUNK_SAMPLES = 480000


def piano_chord(k: int = 0, n_samples: int = CHUNK_SAMPLES, sr: int = SR) -> np.ndarray:
    """Synthetic chunk ``k`` (float32, ``n_samples``): four decaying harmonic
    notes re-struck every 2 s plus a 1e-3 noise floor."""
    rng = np.random.default_rng(k)
    t = np.arange(n_samples, dtype=np.float64) / sr
    env = np.exp(-1.5 * np.mod(t, 2.0))
    y = np.zeros(n_samples, dtype=np.float64)
    for midi in (60, 64, 67, 72):
        f0 = 440.0 * 2.0 ** ((midi + (k % 12) - 69) / 12.0)
        for h in range(1, 6):
            y += (0.5 / h) * np.sin(2.0 * np.pi * f0 * h * t) * env
    y = 0.2 * y + 1e-3 * rng.standard_normal(n_samples)
    return y.astype(np.float32)


def piano_chord_batch(ks, n_samples: int = CHUNK_SAMPLES) -> np.ndarray:
    return np.stack([piano_chord(int(k), n_samples) for k in ks])


def piano_chord_batch_fast(ks, n_samples: int = CHUNK_SAMPLES, sr: int = SR) -> torch.Tensor:
    """``piano_chord`` for many chunks, evaluated with torch's multi-threaded float64 kernels in the same operation
    order (the float32 results are bit-identical to ``piano_chord``, tests/test_packing.py): 240 chunks in ~15 s
    instead of 70.  Returns a float32 CPU tensor (len(ks), n_samples)."""
    ks = [int(k) for k in ks]
    out = torch.empty(len(ks), n_samples, dtype=torch.float32)
    t = torch.arange(n_samples, dtype=torch.float64) / sr
    env = torch.exp(-1.5 * torch.remainder(t, 2.0))
    for i, k in enumerate(ks):
        rng = np.random.default_rng(k)
        y = torch.zeros(n_samples, dtype=torch.float64)
        for midi in (60, 64, 67, 72):
            f0 = 440.0 * 2.0 ** ((midi + (k % 12) - 69) / 12.0)
            for h in range(1, 6):
                y += (0.5 / h) * torch.sin(2.0 * np.pi * f0 * h * t) * env
        out[i] = (0.2 * y + 1e-3 * torch.from_numpy(rng.standard_normal(n_samples))).float()
    return out


def to_pcm16(wav: torch.Tensor) -> torch.Tensor:
    """float waveform in [-1, 1) -> the int16 samples a 16-bit WAVE file of it holds (round to nearest, saturate)."""
    return torch.clamp(torch.round(wav * 32768.0), -32768, 32767).to(torch.int16)


def cheap_wave_batch(n: int, n_samples: int = CHUNK_SAMPLES, seed: int = 0) -> torch.Tensor:
    """Fast batch generator for bench.py (tones + noise, float32 CPU tensor).
    Same spectral character as ``piano_chord`` but vectorised in torch so that
    240 chunks are built in about a second."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float32) / SR
    env = torch.exp(-1.5 * torch.remainder(t, 2.0))
    out = torch.empty(n, n_samples, dtype=torch.float32)
    for i in range(n):
        y = torch.zeros(n_samples)
        for midi in (60, 64, 67, 72):
            f0 = 440.0 * 2.0 ** ((midi + (i % 12) - 69) / 12.0)
            for h in range(1, 4):
                y += (0.5 / h) * torch.sin(2.0 * np.pi * f0 * h * t) * env
        out[i] = 0.2 * y + 1e-3 * torch.randn(n_samples, generator=g)
    return out


# --------------------------------------------------------------------------
# checkpoint key list
# --------------------------------------------------------------------------
def _bn(prefix: str, c: int):
    return [
        (prefix + ".weight", (c,), "bn_w"),
        (prefix + ".bias", (c,), "bn_b"),
        (prefix + ".running_mean", (c,), "bn_m"),
        (prefix + ".running_var", (c,), "bn_v"),
        (prefix + ".num_batches_tracked", (), "bn_n"),
    ]


def _conv(prefix: str, co: int, ci: int, kh: int, kw: int):
    return [(prefix + ".weight", (co, ci, kh, kw), "w"), (prefix + ".bias", (co,), "b")]


def _lstm(prefix: str, inp: int, hid: int, layers: int):
    out = []
    for l in range(layers):
        for suf in ("", "_reverse"):
            i = inp if l == 0 else 2 * hid
            out += [
                (f"{prefix}.weight_ih_l{l}{suf}", (4 * hid, i), "w"),
                (f"{prefix}.weight_hh_l{l}{suf}", (4 * hid, hid), "w"),
                (f"{prefix}.bias_ih_l{l}{suf}", (4 * hid,), "b"),
                (f"{prefix}.bias_hh_l{l}{suf}", (4 * hid,), "b"),
            ]
    return out


def _lin(prefix: str, o: int, i: int):
    return [(prefix + ".weight", (o, i), "w"), (prefix + ".bias", (o,), "b")]


def state_dict_spec(model_type: str, n_mels: int, hidden_size: int, num_layers: int,
                    use_attention: bool = True, use_onset_offset_heads: bool = True):
    """[(key, shape, kind)] in the order ``TranscriptionModel.state_dict()``
    of the reference emits them (keys carry the wrapper's ``model.`` prefix,
    reference models/transcription_model.py:45-59)."""
    mt = model_type.lower()
    H = hidden_size
    spec = []
    if mt in ("cnn_rnn", "cnn+rnn"):
        spec += _conv("model.cnn.0", 32, 1, 3, 3) + _bn("model.cnn.1", 32)
        spec += _conv("model.cnn.4", 64, 32, 3, 3) + _bn("model.cnn.5", 64)
        spec += _lstm("model.rnn", 64 * (n_mels // 4), H, num_layers)
        spec += _lin("model.fc", 88, 2 * H)
    elif mt in ("cnn_rnn_large", "large"):
        spec += _conv("model.conv1.0", 32, 1, 3, 3) + _bn("model.conv1.1", 32)
        for name, ci, co in (("model.res_block1", 32, 64), ("model.res_block2", 64, 128)):
            spec += _conv(name + ".conv1", co, ci, 3, 3) + _bn(name + ".bn1", co)
            spec += _conv(name + ".conv2", co, co, 3, 3) + _bn(name + ".bn2", co)
            spec += _conv(name + ".skip.0", co, ci, 1, 1) + _bn(name + ".skip.1", co)
        spec += _conv("model.freq_aware_conv.0", 256, 128, 7, 3) + _bn("model.freq_aware_conv.1", 256)
        inp = 256 * (n_mels // 8)
        spec += _lstm("model.rnn_main", inp, H, num_layers)
        spec += _lstm("model.rnn_local", inp, H // 2, 1)
        D = 2 * H + 2 * (H // 2)
        if use_attention:
            spec += _lin("model.attention.qkv", 3 * D, D) + _lin("model.attention.proj", D, D)
            spec += [("model.attention_norm.weight", (D,), "ln_w"), ("model.attention_norm.bias", (D,), "bn_b")]
        if use_onset_offset_heads:
            spec += _lin("model.shared_fc", H, D)
            spec += _lin("model.frame_head", 88, H) + _lin("model.onset_head", 88, H) + _lin("model.offset_head", 88, H)
        else:
            spec += _lin("model.fc", 88, D)
    else:
        raise ValueError(f"Unknown model type: {model_type}")
    return spec


def synth_state_dict(model_type: str, n_mels: int, hidden_size: int, num_layers: int,
                     seed: int = 1, use_attention: bool = True,
                     use_onset_offset_heads: bool = True, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded fp32 checkpoint with the reference's key set.  Each tensor has its
    own generator seeded by (seed, crc32(key)), so values do not depend on
    construction order.  Weights ~ U(-a, a) with a = gain*sqrt(3/fan_in)
    (unit-variance preserving), BN stats perturbed as SURVEY.md section 8(d)."""
    sd = OrderedDict()
    for key, shape, kind in state_dict_spec(model_type, n_mels, hidden_size, num_layers,
                                            use_attention, use_onset_offset_heads):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
        if kind == "w":
            fan_in = int(np.prod(shape[1:]))
            a = gain * (3.0 / fan_in) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind in ("bn_w", "ln_w"):
            t = torch.rand(shape, generator=g) + 0.5
        elif kind == "bn_b":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_m":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_v":
            t = torch.rand(shape, generator=g) + 0.5
        elif kind == "bn_n":
            t = torch.tensor(100, dtype=torch.int64)
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[key] = t
    return sd


def synth_logmel(B: int, n_mels: int, T: int, seed: int = 0) -> torch.Tensor:
    """dB-like model input (B,1,n_mels,T) float32 in about [-55, 25]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, n_mels, T, generator=g) * 12.0 - 20.0
    return x.clamp_(-55.0, 26.0)


def planted_probs(n_pitch: int, T: int, thresholds, seed: int = 0, frac: float = 0.01) -> np.ndarray:
    """U(0,1) float32 roll with ``frac`` of the cells overwritten by exact
    float32(threshold) values (strict ``>`` must treat them as inactive)."""
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(n_pitch, T, generator=g).numpy().copy()
    thr = np.asarray(thresholds, dtype=np.float64).astype(np.float32)
    rng = np.random.default_rng(seed + 7)
    n = int(frac * p.size)
    if n and thr.size:
        idx = rng.choice(p.size, size=n, replace=False)
        p.reshape(-1)[idx] = thr[rng.integers(0, thr.size, size=n)]
    return p


def bernoulli_roll(n_pitch: int, T: int, p: float = 0.05, seed: int = 0) -> np.ndarray:
    g = torch.Generator().manual_seed(seed + 100003)
    return (torch.rand(n_pitch, T, generator=g) < p).float().numpy()
