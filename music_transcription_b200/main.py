"""Drop-in for the reference's ``main.py`` (the inference script, reference main.py:27-287): the same module-level
constants and the same function names, positional arguments, defaults and return values,

    load_model(model_path, device="cpu")                                   main.py:27-57
    split_audio_into_chunks(audio_path, chunk_length, sr) -> (chunks, duration)   main.py:60-100
    audio_to_mel(audio_chunk, sr, n_mels, hop_length)                      main.py:103-130
    predict_chunk(model, mel_tensor, device, threshold)                    main.py:133-161
    combine_piano_rolls(piano_rolls, chunk_length, sr, hop_length)         main.py:164-186
    pianoroll_to_midi(pianoroll, fs, min_midi=21)                          main.py:189-226
    transcribe_audio(audio_path, model_path, output_path=None, device=None) -> output path   main.py:229-287
    main()                                                                 main.py:290-362

so that ``import main`` can be replaced by ``from music_transcription_b200 import main``.  Underneath, everything is
the sm_100a library: there is NO CPU path -- ``device="cpu"`` (the reference's default) raises instead of silently
running somewhere else.  ``transcribe_audio`` batches the chunks (main.py loops one by one) and keeps the audio on the
device from decode to note list; its output file holds the same notes.
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

from . import _lib, audio, pipeline
from .pipeline import audio_to_mel, combine_piano_rolls, pianoroll_to_midi, predict_chunk  # noqa: F401  (re-exported)
from .transcription_model import TranscriptionModel

# reference main.py:16-24
MODEL_TYPE = "cnn_rnn_large"
N_MELS = 320
HIDDEN_SIZE = 512
NUM_LAYERS = 3
DROPOUT = 0.2
SR = 16000
HOP_LENGTH = 512
CHUNK_LENGTH = 30.0
THRESHOLD = 0.5


def _cuda_device(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.AmtError(f"device {device!r}: this implementation runs on a B200 only (no CPU fallback); pass device='cuda'")
    return dev


def load_model(model_path, device="cpu"):
    """Load a reference ``.pth`` state_dict (as scripts/train_cnn.py:358 saves it) into the B200 model, eval mode."""
    dev = _cuda_device(device)
    print(f"Loading model from {model_path}...")
    model = TranscriptionModel(model_type=MODEL_TYPE, n_mels=N_MELS, hidden_size=HIDDEN_SIZE, num_layers=NUM_LAYERS,
                               dropout=DROPOUT, device=str(dev))
    checkpoint = torch.load(model_path, map_location=str(dev))
    model.load_state_dict(checkpoint)
    model.eval()
    model.to(dev)
    print(f"Model loaded successfully on {device}")
    return model


def split_audio_into_chunks(audio_path, chunk_length=CHUNK_LENGTH, sr=SR):
    """Decode + resample the file (``audio.load_audio``: what ``librosa.load(path, sr=sr, mono=True)`` does, on the
    GPU) and cut it into ``chunk_length``-second chunks, the last one zero-padded.  Returns (list of float32 numpy
    chunks, duration in seconds) like the reference; ``transcribe_audio`` uses the device-resident form instead."""
    print(f"Loading audio from {audio_path}...")
    y, _ = audio.load_audio(str(audio_path), sr)
    duration = y.numel() / sr
    print(f"Audio duration: {duration:.2f} seconds")
    chunks = [c for c in pipeline.split_audio_into_chunks(y, chunk_length, sr).cpu().numpy()]
    print(f"Split audio into {len(chunks)} chunks of {chunk_length}s each")
    return chunks, duration


def transcribe_audio(audio_path, model_path, output_path=None, device=None):
    """Full pipeline, file to MIDI file (main.py:229-287).  Returns the output path."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    print(f"Using device: {device}")
    model = load_model(model_path, device)
    print("Processing chunks and running predictions...")
    out = audio.transcribe_audio(audio_path, model, output_path, threshold=THRESHOLD)
    print(f"MIDI file saved to: {out}")
    return out


def main():
    """Command line of the reference script: ``main.py audio_file model_file [-o OUT] [-d cuda] [-t THRESHOLD]``."""
    parser = argparse.ArgumentParser(description="Transcribe audio files to MIDI using trained music transcription model")
    parser.add_argument("audio_file", type=str, help="Path to input audio file (wav)")
    parser.add_argument("model_file", type=str, help="Path to model checkpoint file (.pth)")
    parser.add_argument("-o", "--output", type=str, default=None, help="Path to output MIDI file (default: <audio_name>_transcription.mid)")
    parser.add_argument("-d", "--device", type=str, choices=["cpu", "cuda"], default=None, help="Device (default: auto-detect; only cuda runs)")
    parser.add_argument("-t", "--threshold", type=float, default=0.5, help="Threshold for note predictions (default: 0.5)")
    args = parser.parse_args()
    global THRESHOLD
    THRESHOLD = args.threshold
    for what, p in (("Audio", args.audio_file), ("Model", args.model_file)):
        if not os.path.exists(p):
            print(f"Error: {what} file not found: {p}")
            sys.exit(1)
    try:
        output_path = transcribe_audio(args.audio_file, args.model_file, args.output, args.device)
        print(f"Output: {output_path}")
    except Exception as e:                                    # the reference prints and exits 1 (main.py:356-362)
        print(f"Error during transcription: {e}")
        import traceback
        traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
