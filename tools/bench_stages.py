"""Stage-level measurements outside bench.py's headline: BASELINE configs[1] (log-mel frontend alone on 64
synthetic 30-s chunks) and the 8f rank-2 resampler (2 h of 44.1 kHz audio -> 16 kHz), CUDA-event timed,
algorithmic bytes against the measured HBM peak.  One JSON line per stage.

    python tools/bench_stages.py [--chunks 64] [--iters 20]
"""
import argparse, json, os, sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from music_transcription_b200 import audio, pipeline, synth  # noqa: E402


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.add_(1.0)                                   # > L2: evicts the previous iteration's data
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    dev = torch.device("cuda", 0)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)     # 256 MB
    C = args.chunks
    wav = synth.cheap_wave_batch(8, 480000, seed=0).repeat((C + 7) // 8, 1)[:C].to(dev)
    fe = pipeline.Frontend.get(device=dev)
    ms = timed(lambda: fe.logmel(wav), args.iters, flush)
    nbytes = C * (480000 * 4 + 320 * 938 * 4)
    print(json.dumps({"stage": "logmel frontend alone (BASELINE configs[1])", "chunks": C, "ms": round(ms, 4),
                      "chunks_per_s": round(C / ms * 1e3, 1), "algorithmic_GBps": round(nbytes / ms / 1e6, 1),
                      "hbm_peak_GBps": hbm, "frac_hbm": round(nbytes / ms / 1e6 / hbm, 4),
                      "fp32_GFLOP": round(C * 60e-3, 2), "fp32_TFLOPs": round(C * 60e6 / ms / 1e9, 2),
                      "bound": "fp32 issue (FFT butterflies + banded mel), not HBM: see DESIGN.md section 4"}))
    n_in = 44100 * 7200
    x = torch.randn(n_in, device=dev)
    ms = timed(lambda: audio.resample(x, 44100, 16000), max(args.iters // 4, 3), flush)
    n_out = -(-n_in * 160 // 441)
    nbytes = (n_in + n_out) * 4
    print(json.dumps({"stage": "polyphase resampler 44.1 kHz -> 16 kHz, 2 h mono (8f rank 2)", "ms": round(ms, 3),
                      "algorithmic_GBps": round(nbytes / ms / 1e6, 1), "hbm_peak_GBps": hbm,
                      "frac_hbm": round(nbytes / ms / 1e6 / hbm, 4), "taps_per_output": 8821 // 160 + 1}))


if __name__ == "__main__":
    main()
