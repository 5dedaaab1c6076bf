"""Host-side mirror of the reference's inference functions (main.py) on top of
the C ABI.  Same names, argument meaning and return shapes as

    audio_to_mel          main.py:103-130
    predict_chunk         main.py:133-161
    combine_piano_rolls   main.py:164-186
    pianoroll_to_midi     main.py:189-226   (returns a NoteList instead of a pretty_midi object:
                                             pretty_midi is not available offline; the note
                                             fields and their order are identical)

plus the batched entry points the B200 design adds (``Frontend.logmel``,
``extract_notes``, ``transcribe_chunks``): the reference loops over chunks one at
a time with a host<->device hop each (main.py:258-266); here a whole batch of
30-s chunks goes through each kernel at once.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import _lib

# reference defaults (main.py:16-24)
MODEL_TYPE = "cnn_rnn_large"
N_MELS = 320
HIDDEN_SIZE = 512
NUM_LAYERS = 3
DROPOUT = 0.2
SR = 16000
HOP_LENGTH = 512
N_FFT = 2048
CHUNK_LENGTH = 30.0
THRESHOLD = 0.5


class DeferredLogMel:
    """What ``Frontend.logmel(..., defer_floor=True)`` returns: the dB spectrogram BEFORE power_to_db's per-chunk floor
    (reference main.py:125, ``max - top_db``) plus the per-chunk maxima.  ``TranscriptionModel.forward`` accepts it in
    place of the tensor and applies the floor inside its stem convolution's load (amt_model_forward_db) -- bitwise the
    result of flooring first, without reading and re-writing the 1.2 MB per chunk once more.  ``floored()`` gives the
    tensor the reference's ``audio_to_mel`` would have produced."""
    __slots__ = ("mel", "chunk_max", "top_db")

    def __init__(self, mel: torch.Tensor, chunk_max: torch.Tensor, top_db: float):
        self.mel, self.chunk_max, self.top_db = mel, chunk_max, float(top_db)

    @property
    def shape(self):
        return self.mel.shape

    @property
    def device(self):
        return self.mel.device

    def floored(self) -> torch.Tensor:
        return torch.maximum(self.mel, (self.chunk_max - self.top_db).view(-1, 1, 1, 1))


class Frontend:
    """Log-mel frontend handle (filterbank + FFT tables resident on one device)."""

    _cache = {}

    def __init__(self, sr=SR, n_fft=N_FFT, hop_length=HOP_LENGTH, n_mels=N_MELS, fmin=0.0, fmax=None, device=None):
        self.sr, self.n_fft, self.hop, self.n_mels = sr, n_fft, hop_length, n_mels
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise _lib.AmtError("Frontend needs a CUDA device (no CPU fallback)")
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().amt_frontend_create(sr, n_fft, hop_length, n_mels, float(fmin),
                                                      float(fmax if fmax is not None else sr / 2.0), C.byref(h)))
        self._h = h

    @classmethod
    def get(cls, sr=SR, n_mels=N_MELS, hop_length=HOP_LENGTH, device=None) -> "Frontend":
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        key = (sr, n_mels, hop_length, str(dev))
        if key not in cls._cache:
            cls._cache[key] = cls(sr=sr, hop_length=hop_length, n_mels=n_mels, device=dev)
        return cls._cache[key]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().amt_frontend_destroy(self._h)
        except Exception:
            pass

    def num_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop

    def filterbank(self) -> np.ndarray:
        fb = np.empty((self.n_mels, 1 + self.n_fft // 2), dtype=np.float32)
        _lib.check(_lib.lib().amt_frontend_filterbank_host(self._h, fb.ctypes.data))
        return fb

    def logmel(self, wav: torch.Tensor, top_db: float = 80.0, defer_floor: bool = False):
        """wav (B, n_samples) float32 CUDA -> (B, 1, n_mels, T) float32 dB, floor at per-chunk max - top_db.
        ``defer_floor``: return a ``DeferredLogMel`` (unfloored dB + per-chunk maxima) for ``TranscriptionModel.forward``
        to floor while it loads its input -- the fused audio -> notes paths use this."""
        _lib.require_cuda(wav, "Frontend.logmel input")
        if wav.dim() == 1:
            wav = wav[None]
        wav = wav.float()
        if wav.stride(-1) != 1:
            wav = wav.contiguous()
        B, n = wav.shape
        T = self.num_frames(n)
        out = torch.empty(B, 1, self.n_mels, T, dtype=torch.float32, device=wav.device)
        cmax = torch.empty(B, dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            defer = defer_floor and top_db is not None
            _lib.check(_lib.lib().amt_logmel_f32(self._h, _lib.ptr(wav), B, n, wav.stride(0), _lib.ptr(out),
                                                 float(top_db if (top_db is not None and not defer) else -1.0), _lib.ptr(cmax),
                                                 _lib.stream_ptr(wav.device)))
        return DeferredLogMel(out, cmax, top_db) if defer else out


def audio_to_mel(audio_chunk, sr=SR, n_mels=N_MELS, hop_length=HOP_LENGTH, device=None):
    """One chunk of samples (numpy or tensor) -> mel tensor (1, 1, n_mels, T) in dB, on the GPU."""
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    wav = torch.as_tensor(np.asarray(audio_chunk, dtype=np.float32) if not torch.is_tensor(audio_chunk) else audio_chunk)
    wav = wav.to(dev, dtype=torch.float32, non_blocking=True).reshape(1, -1)
    return Frontend.get(sr, n_mels, hop_length, dev).logmel(wav)


def predict_chunk(model, mel_tensor, device, threshold=THRESHOLD):
    """Binary piano roll (88, T) numpy float32 for one mel chunk (1, 1, n_mels, T)."""
    roll = model.predict(mel_tensor.to(device), threshold=threshold)
    return roll[0].cpu().numpy()


def combine_piano_rolls(piano_rolls, chunk_length=CHUNK_LENGTH, sr=SR, hop_length=HOP_LENGTH):
    """np.concatenate along time; a single roll is returned as is (main.py:177-184)."""
    if len(piano_rolls) == 1:
        return piano_rolls[0]
    return np.concatenate(piano_rolls, axis=1)


@dataclass
class Note:
    velocity: int
    pitch: int
    start: float
    end: float


class NoteList:
    """Stand-in for the pretty_midi object main.py builds: ``.instruments[0].notes`` holds
    Note(velocity=100, pitch=21+idx, start=onset/fs, end=offset/fs) in the reference's order."""

    class _Instrument:
        def __init__(self, notes):
            self.program, self.notes = 0, notes

    def __init__(self, triples: np.ndarray, fs: float, min_midi: int = 21):
        self.triples = np.asarray(triples, dtype=np.int32).reshape(-1, 3)
        self.fs, self.min_midi = fs, min_midi
        notes = [Note(100, int(min_midi + p), int(s) / fs, int(e) / fs) for p, s, e in self.triples]
        self.instruments = [NoteList._Instrument(notes)]

    def __len__(self):
        return len(self.triples)

    def write(self, path: str) -> None:
        """``pretty_midi.PrettyMIDI.write`` for this object (main.py:284): a format-1 Standard MIDI File."""
        from . import smf
        smf.write_smf(path, self.instruments[0].notes, self.instruments[0].program)


def extract_notes(vals: torch.Tensor, threshold: float = 0.0, cap: int | None = None) -> np.ndarray:
    """Note grouping on the GPU.  ``vals`` is (88, T) or (n_seg, 88, T) CUDA float32 (probabilities
    or a {0,1} roll); segments are treated as ONE roll concatenated along time, so notes crossing a
    chunk seam merge as in main.py:270-275.  Active iff value > float32(threshold).
    Returns int32 (n_notes, 3): (pitch_idx, onset_frame, offset_frame), pitch-major, onset ascending."""
    _lib.require_cuda(vals, "extract_notes input")
    if vals.dim() == 2:
        vals = vals[None]
    vals = vals.float()
    if vals.stride(-1) != 1:
        vals = vals.contiguous()
    n_seg, n_pitch, T = vals.shape
    dev = vals.device
    if cap is None:
        cap = n_pitch * ((n_seg * T + 1) // 2)          # the most notes a roll of that size can hold
    notes = torch.empty(max(cap, 1), 3, dtype=torch.int32, device=dev)
    counts = torch.empty(n_pitch + 1, dtype=torch.int32, device=dev)
    scratch = torch.empty(2 * n_seg * n_pitch, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().amt_threshold_notes(_lib.ptr(vals), n_seg, n_pitch, T, vals.stride(0), vals.stride(1),
                                                  float(np.float32(threshold)), _lib.ptr(notes), cap, _lib.ptr(counts),
                                                  _lib.ptr(scratch), scratch.numel(), _lib.stream_ptr(dev)))
    total = int(counts[n_pitch].item())
    if total > cap:
        raise _lib.AmtError(f"extract_notes: {total} notes exceed cap {cap}")
    return notes[:total].cpu().numpy()


def extract_notes_async(vals: torch.Tensor, threshold: float, notes_out: torch.Tensor, counts_out: torch.Tensor,
                        scratch: torch.Tensor | None = None) -> None:
    """Launch-only variant of ``extract_notes`` (no host sync): writes int32 triples into ``notes_out``
    (cap = notes_out.shape[0]) and per-pitch counts + total into ``counts_out`` (n_pitch + 1).  ``scratch``:
    int32 device tensor of at least 2 * n_seg * n_pitch elements (allocated from torch's caching allocator when None)."""
    if vals.dim() == 2:
        vals = vals[None]
    n_seg, n_pitch, T = vals.shape
    if scratch is None:
        scratch = torch.empty(2 * n_seg * n_pitch, dtype=torch.int32, device=vals.device)
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().amt_threshold_notes(_lib.ptr(vals), n_seg, n_pitch, T, vals.stride(0), vals.stride(1),
                                                  float(np.float32(threshold)), _lib.ptr(notes_out), notes_out.shape[0],
                                                  _lib.ptr(counts_out), _lib.ptr(scratch), scratch.numel(),
                                                  _lib.stream_ptr(vals.device)))


def extract_notes_from_bits(bits: torch.Tensor, T: int, cap: int | None = None) -> np.ndarray:
    """``extract_notes`` on bit-packed rolls (``pack_roll``): bits int32 (n_seg, 88, ceil(T/32)) CUDA, the segments
    concatenated along time.  Same output as grouping the float roll the bits came from."""
    _lib.require_cuda(bits, "extract_notes_from_bits input")
    if bits.dim() == 2:
        bits = bits[None]
    bits = bits.contiguous()
    n_seg, n_pitch, words = bits.shape
    if words != (T + 31) // 32:
        raise ValueError(f"extract_notes_from_bits: {words} words per row do not hold T = {T} frames")
    dev = bits.device
    if cap is None:
        cap = n_pitch * ((n_seg * T + 1) // 2)
    notes = torch.empty(max(cap, 1), 3, dtype=torch.int32, device=dev)
    counts = torch.empty(n_pitch + 1, dtype=torch.int32, device=dev)
    scratch = torch.empty(2 * n_seg * n_pitch, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().amt_bits_notes(_lib.ptr(bits), n_seg, n_pitch, T, _lib.ptr(notes), cap, _lib.ptr(counts),
                                             _lib.ptr(scratch), scratch.numel(), _lib.stream_ptr(dev)))
    total = int(counts[n_pitch].item())
    if total > cap:
        raise _lib.AmtError(f"extract_notes_from_bits: {total} notes exceed cap {cap}")
    return notes[:total].cpu().numpy()


def extract_notes_onset_aware(frame: torch.Tensor, onset: torch.Tensor, offset: torch.Tensor | None = None, threshold: float = THRESHOLD,
                              onset_threshold: float = THRESHOLD, offset_threshold: float = THRESHOLD, logits: bool = True,
                              cap: int | None = None) -> np.ndarray:
    """Notes from the three heads of ``model(mel, return_all_heads=True)`` (the reference computes the
    onset / offset heads, cnn_rnn_model.py:333-345, but its inference path never decodes them): ``frame`` / ``onset`` /
    ``offset`` are (n_chunks, 88, T) CUDA tensors of logits (``logits=True``: sigmoid is applied) or probabilities; chunks
    are concatenated along time like main.py:270-275.  A rising edge of ``sigmoid(onset) > onset_threshold`` starts a note;
    it ends where frame and onset are both inactive, where the offset head fires, at the next onset, or at the end
    (``amt_onset_notes``; rule restated in oracle/notes.py).  Returns int32 (n, 3) rows (pitch_idx, onset_frame,
    offset_frame), pitch-major -- the rows ``NoteList`` / ``smf`` take."""
    _lib.require_cuda(frame, "extract_notes_onset_aware input")
    if frame.dim() == 2:
        frame, onset = frame[None], onset[None]
        offset = offset[None] if offset is not None else None
    n, P, T = frame.shape
    L = _lib.lib()
    dev = frame.device
    W = (T + 31) // 32
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr(dev)
        packed = []
        for t, thr in ((frame, threshold), (onset, onset_threshold), (offset, offset_threshold)):
            if t is None:
                packed.append(None)
                continue
            if t.shape != frame.shape:
                raise ValueError("extract_notes_onset_aware: the heads must have the same shape")
            t = t.contiguous().float()
            bits = torch.empty(n, P, W, dtype=torch.int32, device=dev)
            _lib.check(L.amt_pack_roll_u32(_lib.ptr(t), n * P, T, float(thr), int(logits), _lib.ptr(bits), sp))
            packed.append(bits)
        cap = int(cap) if cap is not None else P * ((n * T + 1) // 2)
        notes = torch.empty(max(cap, 1), 3, dtype=torch.int32, device=dev)
        counts = torch.empty(P + 1, dtype=torch.int32, device=dev)
        scratch = torch.empty(max(int(L.amt_onset_notes_scratch_ints(P)), 1), dtype=torch.int32, device=dev)
        _lib.check(L.amt_onset_notes(_lib.ptr(packed[0]), _lib.ptr(packed[1]), _lib.ptr(packed[2]), n, P, T, _lib.ptr(notes), cap,
                                     _lib.ptr(counts), _lib.ptr(scratch), scratch.numel(), sp))
        total = int(counts[P].item())
    if total > cap:
        raise _lib.AmtError(f"extract_notes_onset_aware: {total} notes exceed cap {cap}")
    return notes[:total].cpu().numpy()


def pianoroll_to_midi(pianoroll, fs, min_midi=21) -> NoteList:
    """(88, T) roll (numpy or tensor, values {0,1}) -> notes, grouped on the GPU (main.py:204-223)."""
    roll = torch.as_tensor(np.ascontiguousarray(pianoroll, dtype=np.float32)) if not torch.is_tensor(pianoroll) else pianoroll
    if not roll.is_cuda:
        roll = roll.to(f"cuda:{torch.cuda.current_device()}")
    return NoteList(extract_notes(roll, threshold=0.0), fs, min_midi)


def split_audio_into_chunks(y, chunk_length=CHUNK_LENGTH, sr=SR):
    """Chunking of an already-decoded mono signal (main.py:82-97): ceil(len/chunk) chunks, last one
    zero-padded.  A numpy signal gives the reference's list of arrays; a torch tensor (e.g. the CUDA output of
    ``audio.load_audio``) gives one (n_chunks, chunk_samples) tensor on the same device -- no host round trip."""
    chunk_samples = int(chunk_length * sr)
    if isinstance(y, torch.Tensor):
        n = -(-y.numel() // chunk_samples)
        out = torch.zeros(n * chunk_samples, dtype=y.dtype, device=y.device)
        out[:y.numel()] = y.reshape(-1)
        return out.view(n, chunk_samples)
    n = int(np.ceil(len(y) / chunk_samples))
    out = []
    for i in range(n):
        c = y[i * chunk_samples:min((i + 1) * chunk_samples, len(y))]
        if len(c) < chunk_samples:
            c = np.pad(c, (0, chunk_samples - len(c)), mode="constant")
        out.append(c)
    return out


_LANES = {}


def lane_streams(device) -> list:
    """The two compute streams ("lanes") of a device, created once: the model keeps one workspace PER STREAM (12 GB at 64
    chunks), so everything that runs two batches at a time -- transcribe_chunks, StreamingTranscriber, bench.py -- shares
    these two instead of creating its own."""
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _LANES:
        _LANES[key] = [torch.cuda.Stream(dev) for _ in range(2)]
    return _LANES[key]


@torch.no_grad()
def transcribe_chunks(model, wav: torch.Tensor, threshold: float = THRESHOLD, sr=SR, n_mels=N_MELS,
                      hop_length=HOP_LENGTH, batch: int = 64, return_probs: bool = False, lanes: int = 2):
    """Batched main.py:258-275: wav (n_chunks, n_samples) CUDA -> (notes int32 (n,3), probs or None).
    All chunks go through log-mel -> forward -> sigmoid in batches of ``batch``; notes are grouped on the
    concatenated roll.  With more than one batch and ``lanes=2`` consecutive batches run on two CUDA streams (one
    workspace each, the weights shared), so the latency-bound recurrences of one batch overlap the tensor kernels of the
    other (+15 % on a long recording); every chunk's result is bitwise what one stream computes."""
    _lib.require_cuda(wav, "transcribe_chunks input")
    if lanes not in (1, 2):
        raise ValueError("transcribe_chunks: lanes must be 1 or 2")
    dev = wav.device
    fe = Frontend.get(sr, n_mels, hop_length, dev)
    n = wav.shape[0]
    T = fe.num_frames(wav.shape[1])
    probs = torch.empty(n, 88, T, dtype=torch.float32, device=dev)
    L = _lib.lib()
    starts = list(range(0, n, batch))
    cur = torch.cuda.current_stream(dev)
    streams = [cur]
    if lanes == 2 and len(starts) > 1:
        streams = lane_streams(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        for st in streams:
            st.wait_event(ready)                         # `wav` (and `probs`) as the caller's stream left them
    for j, i in enumerate(starts):
        st = streams[j % len(streams)]
        with torch.cuda.device(dev), torch.cuda.stream(st):
            mel = fe.logmel(wav[i:i + batch], defer_floor=True)
            logits = model(mel)
            _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.0, _lib.ptr(probs[i:i + batch]), 0,
                                               st.cuda_stream))
            if st is not cur:                            # the caching allocator may hand these blocks to another stream later
                mel.mel.record_stream(st)
                mel.chunk_max.record_stream(st)
                logits.record_stream(st)
    if len(streams) > 1:
        for st in streams:
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
    notes = extract_notes(probs, threshold=threshold)
    return notes, (probs if return_probs else None)


def pack_roll(vals: torch.Tensor, threshold: float, apply_sigmoid: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
    """(..., T) float32 CUDA (probabilities, or logits with ``apply_sigmoid``) -> (..., ceil(T/32)) int32 words holding
    bit t%32 of word t/32 = (value > float32(threshold)): the piano roll at 1/32 of the float bytes."""
    _lib.require_cuda(vals, "pack_roll input")
    vals = vals.float().contiguous()
    T = vals.shape[-1]
    words = (T + 31) // 32
    if out is None:
        out = torch.empty(*vals.shape[:-1], words, dtype=torch.int32, device=vals.device)
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().amt_pack_roll_u32(_lib.ptr(vals), vals.numel() // T, T, float(np.float32(threshold)),
                                                int(apply_sigmoid), _lib.ptr(out), _lib.stream_ptr(vals.device)))
    return out


def unpack_roll(bits, T: int) -> np.ndarray:
    """Host-side inverse of ``pack_roll``: (..., ceil(T/32)) int32 words (tensor or array) -> (..., T) float32 {0,1},
    the array ``predict_chunk`` returns (main.py:153-160)."""
    a = bits.cpu().numpy() if torch.is_tensor(bits) else np.asarray(bits)
    a = np.ascontiguousarray(a).view(np.uint32)
    b = np.unpackbits(a.view(np.uint8).reshape(*a.shape[:-1], -1), axis=-1, bitorder="little")
    return b[..., :T].astype(np.float32)


class StreamingTranscriber:
    """Host buffers in, host piano-rolls + note lists out, batch after batch, with the copies of batch i+1
    (pinned host audio -> device) and of batch i-1 (rolls and notes -> pinned host) overlapped with the compute
    of batch i, and -- ``lanes=2`` -- TWO batches computing at once on two streams, so that the latency-bound LSTM
    recurrences of one (which leave 50-100 of the 148 SMs idle) run beside the tensor kernels of the other (-11 % time per
    batch, measured; results are bitwise those of one lane).  This is the end-to-end form of main.py:258-275 for a long
    recording -- every batch still pays its H2D and D2H, they just no longer sit on the critical path.

        st = StreamingTranscriber(model, chunks_per_batch=64, input_format="pcm16")
        for roll_bits, notes in st.run(pinned_batches):
            roll = unpack_roll(roll_bits, st.T)         # (C, 88, T) float32 {0,1}; notes: (n, 3) int32

    ``input_format``: "f32" -- pinned float32 (c, n_samples) batches, what ``librosa.load`` hands main.py:76;
        "pcm16" -- pinned int16 (c, n_samples) mono PCM as a 16-bit WAVE file stores it: half the PCIe bytes, converted
        on the device by ``amt_pcm16_to_mono_f32`` (sample / 32768, bit-identical to the host decode).
    ``roll_format``: "bits" (default) -- the roll leaves the device bit-packed, int32 (c, 88, ceil(T/32)),
        10.6 KB instead of 330 KB per chunk (``unpack_roll`` restores the float array); "f32" -- the float {0,1} roll
        of main.py:153-160 itself.
    ``lanes``: batches in flight on the GPU (1 or 2).  Results come out in input order either way.
    LIFETIME: each yielded (roll, notes) pair is a VIEW into one of ``2 * lanes`` reused pinned slots and is valid until
    the generator is advanced again -- consume or copy it inside the loop body (``copy=True`` yields private copies
    instead, for ``list(st.run(...))``).
    ``notes`` are grouped per batch (frame indices relative to the batch); ``sharding.stitch_notes`` merges
    batches / ranks exactly like grouping the concatenated roll, ``sharding.AsyncRollGather`` does it from the packed rolls."""

    class _Slot:
        pass

    def __init__(self, model, chunks_per_batch: int, n_samples: int = int(CHUNK_LENGTH * SR), threshold: float = THRESHOLD,
                 sr=SR, n_mels=N_MELS, hop_length=HOP_LENGTH, input_format: str = "f32", roll_format: str = "bits",
                 copy: bool = False, lanes: int = 1):
        if input_format not in ("f32", "pcm16") or roll_format not in ("bits", "f32"):
            raise ValueError("StreamingTranscriber: input_format in {'f32','pcm16'}, roll_format in {'bits','f32'}")
        if lanes not in (1, 2):
            raise ValueError("StreamingTranscriber: lanes must be 1 or 2")
        self.model, self.C, self.thr = model, chunks_per_batch, float(threshold)
        self.input_format, self.roll_format, self.copy, self.lanes = input_format, roll_format, copy, lanes
        dev = torch.device(model.device)
        _lib.require_cuda(torch.empty(0, device=dev), "StreamingTranscriber device")
        self.dev = dev
        self.fe = Frontend.get(sr, n_mels, hop_length, dev)
        self.T = self.fe.num_frames(n_samples)
        self.n_samples = n_samples
        C, T = chunks_per_batch, self.T
        W = (T + 31) // 32
        cap = 88 * ((C * T + 1) // 2)
        self.copy_in, self.copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # the note list is fetched in a second step (its length must reach the host first); on its own stream, so that it
        # never queues behind the roll download of a LATER batch that is still computing
        self.copy_notes = torch.cuda.Stream(dev)
        # lanes == 1 computes on the caller's current stream (as before); lanes == 2 on two streams of its own
        self.lane_streams = lane_streams(dev) if lanes > 1 else [None]
        self.slots = []
        for _ in range(2 * lanes):
            s = StreamingTranscriber._Slot()
            s.wav = torch.empty(C, n_samples, device=dev)
            s.pcm = torch.empty(C, n_samples, dtype=torch.int16, device=dev) if input_format == "pcm16" else None
            s.probs = torch.empty(C, 88, T, device=dev)
            if roll_format == "f32":
                s.roll = torch.empty(C, 88, T, device=dev)
                s.host_roll = torch.empty(C, 88, T, dtype=torch.float32).pin_memory()
            else:
                s.roll = torch.empty(C, 88, W, dtype=torch.int32, device=dev)
                s.host_roll = torch.empty(C, 88, W, dtype=torch.int32).pin_memory()
            s.notes = torch.empty(cap, 3, dtype=torch.int32, device=dev)
            s.counts = torch.zeros(89, dtype=torch.int32, device=dev)
            s.scratch = torch.empty(2 * 88 * C, dtype=torch.int32, device=dev)
            s.host_counts = torch.zeros(89, dtype=torch.int32).pin_memory()
            s.host_notes = torch.empty(cap, 3, dtype=torch.int32).pin_memory()
            s.h2d_done, s.compute_done, s.counts_done, s.d2h_done = (torch.cuda.Event() for _ in range(4))
            s.n = C
            self.slots.append(s)
        self.h2d_bytes = C * n_samples * (2 if input_format == "pcm16" else 4)
        self.roll_bytes = s.host_roll.numel() * 4 + 89 * 4

    def _launch(self, i: int, host_wav: torch.Tensor) -> None:
        s = self.slots[i % len(self.slots)]
        lane = self.lane_streams[i % self.lanes]
        compute = lane if lane is not None else torch.cuda.current_stream(self.dev)
        n = host_wav.shape[0]
        want = torch.int16 if self.input_format == "pcm16" else torch.float32
        if host_wav.dtype != want or host_wav.shape[1] != self.n_samples or n > self.C:
            raise ValueError(f"StreamingTranscriber: expected {want} batches of shape (<= {self.C}, {self.n_samples}), "
                             f"got {host_wav.dtype} {tuple(host_wav.shape)}")
        s.n = n
        L = _lib.lib()
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(s.compute_done)             # the compute that last read this slot's audio is done
            (s.pcm if s.pcm is not None else s.wav)[:n].copy_(host_wav, non_blocking=True)
            s.h2d_done.record(self.copy_in)
        with torch.cuda.device(self.dev), torch.cuda.stream(compute):
            compute.wait_event(s.h2d_done)
            compute.wait_event(s.d2h_done)                      # this slot's previous results have left the device
            sp = compute.cuda_stream
            if s.pcm is not None:
                _lib.check(L.amt_pcm16_to_mono_f32(_lib.ptr(s.pcm), n * self.n_samples, 1, _lib.ptr(s.wav), sp))
            mel = self.fe.logmel(s.wav[:n], defer_floor=True)
            logits = self.model(mel)
            if self.roll_format == "f32":
                _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), self.thr, _lib.ptr(s.probs), _lib.ptr(s.roll), sp))
            else:
                _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), self.thr, _lib.ptr(s.probs), 0, sp))
                _lib.check(L.amt_pack_roll_u32(_lib.ptr(s.probs), n * 88, self.T, self.thr, 0, _lib.ptr(s.roll), sp))
            extract_notes_async(s.probs[:n], self.thr, s.notes, s.counts, s.scratch)
            s.compute_done.record(compute)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(s.compute_done)
            s.host_roll[:n].copy_(s.roll[:n], non_blocking=True)
            s.host_counts.copy_(s.counts, non_blocking=True)
            s.counts_done.record(self.copy_out)

    def _finish(self, i: int):
        s = self.slots[i % len(self.slots)]
        s.counts_done.synchronize()                             # host waits; the GPU is already on the next batches
        total = int(s.host_counts[88])
        if total > s.notes.shape[0]:
            raise _lib.AmtError(f"StreamingTranscriber: {total} notes exceed the buffer")
        with torch.cuda.stream(self.copy_notes):
            self.copy_notes.wait_event(s.counts_done)           # (implies this batch's compute)
            s.host_notes[:total].copy_(s.notes[:total], non_blocking=True)
            s.d2h_done.record(self.copy_notes)
        s.d2h_done.synchronize()
        roll, notes = s.host_roll[:s.n], s.host_notes[:total].numpy()
        return (roll.clone(), notes.copy()) if self.copy else (roll, notes)

    def run(self, host_batches):
        """host_batches: iterable of pinned (c <= chunks_per_batch, n_samples) tensors, float32 or int16 per
        ``input_format``.  Yields (roll, notes) per batch, in order -- see the class docstring for formats and lifetime."""
        depth = self.lanes                                      # batches launched ahead of the one being collected
        i = -1
        for i, hb in enumerate(host_batches):
            self._launch(i, hb)
            if i >= depth:
                yield self._finish(i - depth)
        for j in range(max(i - depth + 1, 0), i + 1):
            yield self._finish(j)
