"""GPU debugging aid: is a chunk's result independent of its batch position, run to run and stage by stage?
    python tools/debug_batch_invariance.py [B] [chunk]"""
import sys

import torch

sys.path.insert(0, ".")
from music_transcription_b200 import pipeline, synth
from music_transcription_b200.transcription_model import TranscriptionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pick = int(sys.argv[2]) if len(sys.argv) > 2 else 9
H = 512
DEV = "cuda:0"
sd = synth.synth_state_dict("cnn_rnn_large", 320, H, 3, seed=1, gain=3 ** -0.5)
m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=H, num_layers=3, device=DEV)
m.load_state_dict(sd)
wav = torch.from_numpy(synth.piano_chord_batch(range(16))).to(DEV)
if B > 16:
    wav = torch.cat([wav * (1.0 - 0.05 * j) for j in range(B // 16)])
mel = pipeline.Frontend.get(device=DEV).logmel(wav[:B])
T = mel.shape[-1]
D = 3 * H
BUFS = [("act1", torch.bfloat16, 160 * 32), ("h1", torch.bfloat16, 160 * 64), ("act2", torch.bfloat16, 80 * 64),
        ("h2", torch.bfloat16, 80 * 128), ("act3", torch.bfloat16, 80 * 128), ("feat", torch.bfloat16, 40 * 256),
        ("gx", torch.float32, 8 * H), ("seq_a", torch.bfloat16, 2 * H), ("seq_b", torch.bfloat16, 2 * H),
        ("rnn_f32", torch.float32, D), ("qkv", torch.bfloat16, 3 * D), ("att", torch.bfloat16, D), ("proj", torch.float32, D),
        ("normed", torch.bfloat16, D), ("shared", torch.bfloat16, H), ("logits", torch.float32, 384)]


def run(x):
    out = m(x).clone()
    torch.cuda.synchronize()
    b = x.shape[0]
    return out, {n: m.workspace_tensor(n, b, T, dt, w).clone() for n, dt, w in BUFS}


o1, s1 = run(mel)
o2, s2 = run(mel)
print("run-to-run deterministic:", torch.equal(o1, o2))
for n, _, _ in BUFS:
    if not torch.equal(s1[n], s2[n]):
        print("  nondeterministic stage:", n, (s1[n].float() - s2[n].float()).abs().max().item())
for i in sorted(set([0, 5, pick, B - 1])):
    os_, ss = run(mel[i:i + 1])
    line = [f"chunk {i}: logits equal={torch.equal(os_[0], o1[i])}"]
    for n, _, _ in BUFS:
        d = (ss[n][0].float() - s1[n][i].float()).abs().max().item()
        if d != 0:
            line.append(f"{n}:{d:.2e}")
    print(" ".join(line))
