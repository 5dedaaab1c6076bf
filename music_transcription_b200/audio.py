"""Audio loading in front of the hot path: what
``librosa.load(audio_path, sr=16000, mono=True)`` does at reference main.py:76 -- decode, mix to mono,
resample to 16 kHz -- followed by the reference's own chunking (main.py:82-97, ``pipeline.
split_audio_into_chunks``) and, in ``transcribe_audio``, the whole of main.py:229-287.

* decode: RIFF/WAVE PCM (8/16/24/32-bit integer, 32/64-bit float, WAVE_FORMAT_EXTENSIBLE) parsed here;
  other containers (mp3, flac, ...) need the decoders librosa delegates to (soundfile / audioread), which
  are third-party and absent offline -- they raise ``ValueError``.
* mono: mean over channels (``librosa.to_mono``).
* resample: polyphase Kaiser-windowed-sinc FIR on the GPU (``amt_resample_poly_f32``).  librosa's default
  ``res_type='soxr_hq'`` lives in the third-party ``soxr`` library (absent, algorithm not restated);
  the filter here is the one ``scipy.signal.resample_poly`` designs (librosa's ``res_type='polyphase'``):
  ``firwin(2*10*max(up,down)+1, 1/max(up,down), window=('kaiser', 5.0)) * up``.
  **Parity unpinned** against soxr; pinned against scipy in tests/test_audio.py and the GPU suite.
"""
from __future__ import annotations

import math
import struct
from pathlib import Path

import numpy as np
import torch

from . import _lib, pipeline

SR = pipeline.SR


# ----------------------------------------------------------------------------- decode
def _parse_wav(path: str):
    """RIFF/WAVE container -> (format tag, channels, sample rate, bits per sample, raw data bytes)."""
    data = Path(path).read_bytes()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file (other containers need soundfile/audioread, absent here)")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:                 # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    return fmt + (pcm,)


def load_wav_pcm16(path: str):
    """16-bit PCM WAVE -> (int16 array [n_frames, n_channels] exactly as stored, sample rate); None for any other
    sample format (use load_wav).  The raw samples go to the GPU as they are (``load_audio``)."""
    tag, ch, sr, bits, pcm = _parse_wav(path)
    if tag != 1 or bits != 16:
        return None
    x = np.frombuffer(pcm[:len(pcm) // 2 * 2], "<i2")
    n = len(x) // ch
    return x[:n * ch].reshape(n, ch), int(sr)


def load_wav(path: str):
    """RIFF/WAVE -> (float32 array [n_frames, n_channels] in [-1, 1), sample rate)."""
    tag, ch, sr, bits, pcm = _parse_wav(path)
    if tag == 1:                                                   # integer PCM
        if bits == 8:
            x = (np.frombuffer(pcm, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(pcm[:len(pcm) // 2 * 2], "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(pcm[:len(pcm) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(pcm[:len(pcm) // 4 * 4], "<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:                                                 # IEEE float
        x = np.frombuffer(pcm, "<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    n = len(x) // ch
    return x[:n * ch].reshape(n, ch), int(sr)


# ----------------------------------------------------------------------------- resample
def polyphase_taps(up: int, down: int) -> np.ndarray:
    """The filter scipy.signal.resample_poly(x, up, down) designs by default, times ``up`` (float64)."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    m = np.arange(-half_len, half_len + 1, dtype=np.float64)
    f_c = 1.0 / max_rate
    h = f_c * np.sinc(f_c * m) * np.kaiser(2 * half_len + 1, 5.0)
    return h / h.sum() * up


_TAPS_CACHE = {}


def _device_taps(up: int, down: int, device) -> torch.Tensor:
    key = (up, down, str(device))
    if key not in _TAPS_CACHE:
        _TAPS_CACHE[key] = torch.from_numpy(polyphase_taps(up, down).astype(np.float32)).to(device)
    return _TAPS_CACHE[key]


def resample(y, orig_sr: int, target_sr: int, device="cuda") -> torch.Tensor:
    """1-D float32 signal at orig_sr -> CUDA float32 tensor at target_sr (ceil(n * target / orig) samples)."""
    y = torch.as_tensor(y, dtype=torch.float32).to(device).contiguous()
    _lib.require_cuda(y, "resample input")
    if y.dim() != 1:
        raise ValueError("resample expects a 1-D signal")
    if orig_sr == target_sr or y.numel() == 0:
        return y
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    n_out = -(-y.numel() * up // down)
    taps = _device_taps(up, down, y.device)
    out = torch.empty(n_out, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().amt_resample_poly_f32(_lib.ptr(y), y.numel(), _lib.ptr(out), n_out, _lib.ptr(taps),
                                                    taps.numel(), up, down, _lib.stream_ptr(y.device)))
    return out


def load_audio(path: str, sr: int = SR, mono: bool = True, device="cuda"):
    """``librosa.load(path, sr=sr, mono=mono)`` for WAVE files -> (CUDA float32 tensor, sr)."""
    if not mono:
        raise ValueError("the reference always loads mono (main.py:76)")
    raw = load_wav_pcm16(path)
    if raw is not None and raw[0].shape[0] > 0:
        # the common case: 16-bit PCM goes up as int16 (half the PCIe bytes, no host conversion pass) and becomes
        # mono float32 on the GPU, bit-identical to the host path below
        pcm, file_sr = raw
        dev_pcm = torch.from_numpy(np.ascontiguousarray(pcm)).to(device, non_blocking=True)
        y = torch.empty(pcm.shape[0], dtype=torch.float32, device=dev_pcm.device)
        with torch.cuda.device(dev_pcm.device):
            _lib.check(_lib.lib().amt_pcm16_to_mono_f32(_lib.ptr(dev_pcm), pcm.shape[0], pcm.shape[1], _lib.ptr(y),
                                                        _lib.stream_ptr(dev_pcm.device)))
        return resample(y, file_sr, sr, device), sr
    x, file_sr = load_wav(path)
    y = x.mean(axis=1, dtype=np.float32) if x.shape[1] > 1 else x[:, 0]
    return resample(y, file_sr, sr, device), sr


# ----------------------------------------------------------------------------- main.py:229-287
def transcribe_audio(audio_path, model, output_path=None, threshold=pipeline.THRESHOLD, batch: int = 64):
    """Reference ``transcribe_audio`` with an already constructed ``TranscriptionModel``: load -> 30-s
    chunks (last one zero padded) -> batched log-mel / forward / sigmoid -> notes grouped on the concatenated
    roll -> Standard MIDI File next to the input (``<stem>_transcription.mid``) unless ``output_path``."""
    y, _ = load_audio(str(audio_path), SR, device=model.device)
    wav = pipeline.split_audio_into_chunks(y, pipeline.CHUNK_LENGTH, SR)      # stays on the device: (n_chunks, 480000)
    triples, _ = pipeline.transcribe_chunks(model, wav, threshold=threshold, batch=batch)
    midi = pipeline.NoteList(triples, SR / pipeline.HOP_LENGTH, 21)
    if output_path is None:
        p = Path(audio_path)
        output_path = p.parent / f"{p.stem}_transcription.mid"
    midi.write(str(output_path))
    return output_path
