"""GPU soak test of the two-lane mode (two forwards in flight on two streams): runs for SECONDS (one model object, one workspace per stream), checks progress every
100 batches with a watchdog (no progress for 10 s = hang -> exit 3) and reports the ms/batch distribution.
    python tools/soak_two_lanes.py CHUNKS SECONDS [rec_priority|none] [lanes]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from music_transcription_b200 import pipeline, synth
from music_transcription_b200.transcription_model import TranscriptionModel

C = int(sys.argv[1]) if len(sys.argv) > 1 else 60
SECONDS = float(sys.argv[2]) if len(sys.argv) > 2 else 60
prio = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "none" else None
NL = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5)
fe = pipeline.Frontend.get(device=dev)
base = synth.cheap_wave_batch(8, 480000, seed=0)
wav = torch.stack([base[i % 8] for i in range(C)]).to(dev)
lanes = []
for _ in range(NL):
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, device=dev).eval()
    m.load_state_dict(sd)
    s = torch.cuda.Stream(dev, priority=prio) if prio is not None else torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        ref = m(fe.logmel(wav)).clone()
    lanes.append((m, s))
torch.cuda.synchronize()
t_start, n, times, bad = time.time(), 0, [], 0
while time.time() - t_start < SECONDS:
    t0 = time.time()
    outs = []
    for i in range(100):
        m, s = lanes[i % NL]
        with torch.cuda.stream(s):
            outs.append(m(fe.logmel(wav)))
    evs = []
    for _, s in lanes:
        e = torch.cuda.Event()
        e.record(s)
        evs.append(e)
    while not all(e.query() for e in evs):
        if time.time() - t0 > 10 + 0.05 * 100:
            print(f"HANG after {n} batches ({time.time() - t_start:.0f} s)", flush=True)
            os._exit(3)
        time.sleep(0.002)
    times.append(1e3 * (time.time() - t0) / 100)
    bad += int(not torch.equal(outs[-1], ref)) + int(not torch.equal(outs[-2], ref))
    n += 100
t = np.array(times)
print(f"soak ok: lanes {NL} chunks {C} prio {prio}: {n} batches in {time.time() - t_start:.0f} s, ms/batch median {np.median(t):.2f} "
      f"p5 {np.percentile(t, 5):.2f} p95 {np.percentile(t, 95):.2f} max {t.max():.2f}; wrong results: {bad}", flush=True)
