"""Diagnostic: where does the canonical-config GPU output differ from the bf16 emulation / oracle?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from music_transcription_b200 import synth, packing, pipeline
from music_transcription_b200.transcription_model import TranscriptionModel
from oracle import model as omodel
from tests.emulate import emu_forward

H = int(os.environ.get("DBG_H", 512)); L = int(os.environ.get("DBG_L", 3)); nm = int(os.environ.get("DBG_M", 320))
nchunk = int(os.environ.get("DBG_B", 2)); ns = int(os.environ.get("DBG_NS", 480000))
dev = "cuda:0"
sd = synth.synth_state_dict("cnn_rnn_large", nm, H, L, seed=1)
m = TranscriptionModel("cnn_rnn_large", n_mels=nm, hidden_size=H, num_layers=L, device=dev); m.load_state_dict(sd)
wav = torch.from_numpy(synth.piano_chord_batch(range(nchunk), n_samples=ns)).to(dev)
fe = pipeline.Frontend(n_mels=nm, device=dev)
mel = fe.logmel(wav)
o1 = m(mel); o2 = m(mel)
print("deterministic:", torch.equal(o1, o2), (o1 - o2).abs().max().item())
solo = torch.cat([m(mel[i:i+1]) for i in range(nchunk)])
print("batch-vs-solo max diff:", (solo - o1).abs().max().item())
torch.set_num_threads(os.cpu_count())
t = time.time(); ref = omodel.large_forward(sd, mel.cpu(), H, L); print("oracle", time.time() - t)
P = packing.pack_state_dict(sd, "cnn_rnn_large", nm, H, L)
t = time.time(); emu = emu_forward(P, mel.cpu(), "cnn_rnn_large", nm, H, L)["frame"]; print("emu", time.time() - t)
g = o1.cpu()
for name, r in (("oracle", ref), ("emu", emu)):
    d = (g - r).abs()
    print(name, "max", d.max().item(), "mean", d.mean().item(), "| emu-vs-oracle", (emu - ref).abs().max().item())
    T = d.shape[-1]
    for b in range(nchunk):
        bt = d[b].max(0).values
        edges = [0, 8, 64, 128, 256, 512, 768, T - 64, T - 8, T]
        print("  chunk", b, "by time:", [round(bt[a:c].max().item(), 3) for a, c in zip(edges, edges[1:]) if c > a])
        print("  chunk", b, "by pitch max:", round(d[b].max(1).values.max().item(), 3), "worst frames", torch.topk(bt, 5).indices.tolist())
