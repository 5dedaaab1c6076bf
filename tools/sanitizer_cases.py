"""One small invocation of every kernel of libamt_sm100.so, for compute-sanitizer (tools/gpu_sanitizer.sh):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_cases.py [case ...]
Shapes are tiny (the tools slow kernels down 10-1000x) but reach every code path: CTA pairs, the 16-CTA LSTM cluster
and the cooperative fallback, both attention kernels, pooled / skip / split conv epilogues, the weight packer."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from music_transcription_b200 import _lib, evaluate, pipeline, synth
from music_transcription_b200.packing import slice_order, split_act, split_k
from music_transcription_b200.transcription_model import TranscriptionModel

DEV = torch.device("cuda:0")
L = _lib.lib()
S = lambda: _lib.stream_ptr(DEV)
bf = lambda x: x.to(torch.bfloat16)


def gemm():
    for M, N, K, relu, f32 in ((200, 128, 128, 1, 0), (129, 192, 64, 0, 1), (300, 256, 192, 0, 0)):
        a, w, b = bf(torch.randn(M, K)).to(DEV), bf(torch.randn(N, K)).to(DEV), torch.randn(N).to(DEV)
        c = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=DEV)
        _lib.check(L.amt_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(c), M, N, K, N, relu, f32, S()))
    torch.cuda.synchronize()


def conv():
    for B, T, F, Ci, Co, kf, pool, skip, split in ((1, 20, 16, 64, 64, 3, 0, 0, 0), (2, 21, 12, 32, 64, 3, 1, 0, 0), (1, 18, 9, 128, 256, 7, 1, 0, 0),
                                                   (1, 17, 16, 128, 128, 3, 0, 64, 0), (1, 19, 16, 64, 64, 3, 1, 32, 1)):
        x, w = torch.randn(B, T, F, Ci), torch.randn(Co, kf * 3 * Ci) / (Ci * kf * 3) ** 0.5
        x2 = torch.randn(B, T, F, skip) if skip else None
        w2 = torch.randn(Co, skip) if skip else None
        if split:
            xs, ws = split_act(x, Ci), split_k(w, Ci)
            x2s = split_act(x2, skip) if skip else None
            ws = torch.cat([ws, split_k(w2, skip)], 1) if skip else ws
        else:
            xs, ws, x2s = bf(x), bf(torch.cat([w, w2], 1) if skip else w), bf(x2) if skip else None
        xs, ws = xs.to(DEV).contiguous(), ws.to(DEV).contiguous()
        x2s = x2s.to(DEV).contiguous() if skip else None
        Fo = F // 2 if pool else F
        out = torch.empty(B, T, Fo, Co * (3 if split else 1), dtype=torch.bfloat16, device=DEV)
        bias = torch.randn(Co).to(DEV)
        _lib.check(L.amt_conv_bf16(_lib.ptr(xs), _lib.ptr(x2s), _lib.ptr(ws), _lib.ptr(bias), _lib.ptr(out), B, T, F, xs.shape[-1],
                                   x2s.shape[-1] if skip else 0, Co, kf, 3, 1, pool | (2 if split else 0), S()))
    torch.cuda.synchronize()


def lstm():
    for B, T, Hs in ((3, 6, (128, 128)), (5, 5, (512, 512, 256, 256)), (2, 4, (640, 640))):
        seqs = (_lib.LstmSeq * len(Hs))()
        keep = []
        for i, H in enumerate(Hs):
            whh = bf(torch.randn(4 * H, H) / H ** 0.5)[slice_order(H)].contiguous().to(DEV)
            gx = torch.randn(B * T, 4 * H).to(DEV)
            ob, of = torch.empty(B * T, H, dtype=torch.bfloat16, device=DEV), torch.empty(B * T, H, device=DEV)
            keep += [whh, gx, ob, of]
            seqs[i] = _lib.LstmSeq(_lib.ptr(whh), _lib.ptr(gx), _lib.ptr(ob), _lib.ptr(of), H, i & 1, 4 * H, H, H)
        nb = L.amt_lstm_scratch_bytes(seqs, len(Hs), B)
        scratch = torch.empty(nb, dtype=torch.uint8, device=DEV)
        _lib.check(L.amt_lstm_recurrence(seqs, len(Hs), B, T, _lib.ptr(scratch), nb, S()))
        torch.cuda.synchronize()


def attention():
    for B, T, hd in ((1, 70, 64), (2, 150, 192), (1, 70, 48)):
        D = 8 * hd
        qkv = bf(torch.randn(B * T, 3 * D)).to(DEV)
        out = torch.empty(B * T, D, dtype=torch.bfloat16, device=DEV)
        _lib.check(L.amt_attention_bf16(_lib.ptr(qkv), _lib.ptr(out), B, T, 8, hd, 10.0, S()))
    torch.cuda.synchronize()


def integers():
    p = torch.rand(3, 88, 70, device=DEV)
    pipeline.extract_notes(p, 0.5)
    pipeline.unpack_roll(pipeline.pack_roll(p, 0.5), 70)
    y = (torch.rand(3, 88, 70, device=DEV) < 0.1).float()
    evaluate.f1_counts(p, y, [70, 33, 1], np.linspace(0.05, 0.95, 19))
    x = torch.randn(1000, device=DEV)
    pr, ro = torch.empty_like(x), torch.empty_like(x)
    _lib.check(L.amt_sigmoid_threshold(_lib.ptr(x), 1000, 0.5, _lib.ptr(pr), _lib.ptr(ro), S()))
    o = torch.empty(7, 3 * 64, dtype=torch.bfloat16, device=DEV)
    _lib.check(L.amt_split3_bf16(_lib.ptr(torch.randn(7, 64, device=DEV)), 1, _lib.ptr(o), 7, 64, S()))
    torch.cuda.synchronize()


def frontend():
    pipeline.audio_to_mel(synth.piano_chord(0, n_samples=20000), device=DEV)
    pipeline.audio_to_mel(synth.piano_chord(1, n_samples=3000), device=DEV)
    from music_transcription_b200 import audio
    audio.resample(torch.randn(5000), 44100, 16000, DEV)
    pcm = torch.randint(-3000, 3000, (2000, 2), dtype=torch.int16, device=DEV)
    out = torch.empty(2000, device=DEV)
    _lib.check(L.amt_pcm16_to_mono_f32(_lib.ptr(pcm), 2000, 2, _lib.ptr(out), S()))
    torch.cuda.synchronize()


def model():
    for mt, prec in (("cnn_rnn_large", "fast"), ("cnn_rnn_large", "precise"), ("cnn_rnn", "fast")):
        m = TranscriptionModel(mt, n_mels=64, hidden_size=128, num_layers=2, device=DEV, precision=prec)
        m.load_state_dict(synth.synth_state_dict(mt, 64, 128, 2, seed=3))
        m.eval()
        x = synth.synth_logmel(2, 64, 40, seed=1).to(DEV)
        out = m(x, return_all_heads=True)
        m.compute_loss(out, (torch.rand(2, 88, 40, device=DEV) < 0.1).float(), torch.tensor([40, 17]))
        torch.cuda.synchronize()


CASES = {"gemm": gemm, "conv": conv, "lstm": lstm, "attention": attention, "integers": integers, "frontend": frontend, "model": model}
if __name__ == "__main__":
    torch.manual_seed(0)
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
        print("case", name, "ok", flush=True)
