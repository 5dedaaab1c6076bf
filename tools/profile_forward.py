"""Run the CNNRNNModelLarge forward (canonical config) a few times on B chunks of synthetic log-mel: a small target for
`ncu -k regex:<kernel>` captures of one kernel family (per forward: conv1, 5 conv_halo launches -- res1.c1, res1.c2,
res2.c1, res2.c2, freq --, 7 tc_gemm launches -- rnn0/1/2.gemm, attn.qkv, attn.proj, fc1, heads --, 3 lstm_cluster,
attention_tc, add_layernorm, heads_transpose)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_transcription_b200 import synth
from music_transcription_b200.transcription_model import TranscriptionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, device=dev)
m.load_state_dict(synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5))
x = synth.synth_logmel(B, 320, 938, seed=0).to(dev)
for _ in range(n):
    out = m(x, return_all_heads=True)
torch.cuda.synchronize()
print("ok", out["frame"].shape, float(out["frame"].abs().mean()))
