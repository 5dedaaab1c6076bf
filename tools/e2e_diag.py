"""GPU diagnosis: what does a concurrent host<->device copy cost the compute of one batch?
    python tools/e2e_diag.py [chunks]
Prints ms per batch for: compute alone / + concurrent H2D (pinned int16) / + concurrent D2H / the streamer, and the
per-stage times with and without the H2D."""
import sys
import time

import torch

sys.path.insert(0, ".")
from music_transcription_b200 import _lib, pipeline, synth
from music_transcription_b200.transcription_model import TranscriptionModel

C = int(sys.argv[1]) if len(sys.argv) > 1 else 60
K = 10
dev = torch.device("cuda:0")
L = _lib.lib()
sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5)
m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, device=dev)
m.load_state_dict(sd)
fe = pipeline.Frontend.get(device=dev)
base = synth.cheap_wave_batch(8, 480000, seed=0)
wav_host = torch.stack([base[i % 8] * (1 - 0.01 * (i // 8)) for i in range(C)])
pcm_host = synth.to_pcm16(wav_host).pin_memory()
wav = wav_host.to(dev)
pcm_dev = torch.empty_like(pcm_host, device=dev)
probs = torch.empty(C, 88, 938, device=dev)
bits = torch.empty(C, 88, 30, dtype=torch.int32, device=dev)
bits_host = torch.empty(C, 88, 30, dtype=torch.int32).pin_memory()
side = torch.cuda.Stream(dev)
stream = _lib.stream_ptr(dev)


def compute():
    logits = m(fe.logmel(wav))
    _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.5, _lib.ptr(probs), 0, stream))


def timed(fn, k=K):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def with_h2d(pieces=1):
    def f():
        with torch.cuda.stream(side):
            n = C // pieces
            for i in range(pieces):
                pcm_dev[i * n:(i + 1) * n].copy_(pcm_host[i * n:(i + 1) * n], non_blocking=True)
        compute()
    return f


def with_d2h():
    with torch.cuda.stream(side):
        bits_host.copy_(bits, non_blocking=True)
    compute()


print(f"chunks {C}")
print("compute alone          %.3f ms" % timed(compute))
print("H2D alone (int16)      %.3f ms" % timed(lambda: pcm_dev.copy_(pcm_host, non_blocking=True)))
print("+ concurrent H2D       %.3f ms" % timed(with_h2d(1)))
print("+ concurrent H2D x8    %.3f ms" % timed(with_h2d(6)))
print("+ concurrent D2H       %.3f ms" % timed(with_d2h))
print("compute alone again    %.3f ms" % timed(compute))
for fmt_in, fmt_roll in (("pcm16", "bits"), ("f32", "f32")):
    st = pipeline.StreamingTranscriber(m, C, 480000, 0.5, input_format=fmt_in, roll_format=fmt_roll)
    src = pcm_host if fmt_in == "pcm16" else wav_host.pin_memory()

    def run(k):
        for _ in st.run(src for _ in range(k)):
            pass
    run(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(K)
    torch.cuda.synchronize()
    print("streamer %-5s/%-4s     %.3f ms per batch (wall)" % (fmt_in, fmt_roll, 1e3 * (time.perf_counter() - t0) / K))

for label, fn in (("alone", compute), ("with H2D", with_h2d(1))):
    m.profile(True)
    timed(fn, 5)
    st = {n: ms / max(k, 1) for n, ms, k in m.profile_read()}
    m.profile(False)
    print(label, " ".join(f"{n}:{v:.3f}" for n, v in st.items()), "sum %.3f" % sum(st.values()))
