"""Test tooling (CPU, not a pytest file): WHERE does the distance between the bf16 kernels and the fp32 reference
come from?  A torch emulation of the kernels' arithmetic (same layouts as tests/emulate.py) in which every rounding
site can be switched on alone, run at the canonical config (CNNRNNModelLarge 320/512/3, one 30-s chord chunk,
default-init-scale weights) against the fp32 oracle.

    python tests/attribution.py [--chunks 0 3] [--gain 0.577] > profiles/r2_error_attribution.txt

Rounding sites ("bf16" = round to nearest even to 8 mantissa bits, "split" = hi + lo bf16 pair, i.e. 16 bits):
    w_conv   conv weights (res1/res2/freq)           a_conv   conv activations (act1..act3, h1, h2)
    feat     conv-stack output = layer-0 GEMM input  w_ih     LSTM input-projection weights
    w_hh     recurrent weights                        h_fb     h_{t-1} fed back into the recurrent MMA
    seq      inter-layer LSTM outputs                 tanhap   tanh.approx.f32 in the gates (2^-11 rel error model)
    rnn_out  LSTM features into attention / heads     w_attn   qkv / proj weights
    qkv      q, k, v activations                      p_att    softmax probabilities P (bf16 before P.V)
    att      attention output                         normed   LayerNorm output
    w_head   shared_fc / head weights                 shared   shared_fc output
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SITES = ["w_conv", "a_conv", "feat", "w_ih", "w_hh", "h_fb", "seq", "tanhap", "rnn_out", "w_attn", "qkv", "p_att", "att",
         "normed", "w_head", "shared"]


def rnd(x, mode):
    """mode: None/False = exact fp32, 'bf16', 'split' (hi+lo bf16 = 16 mantissa bits)."""
    if not mode:
        return x
    hi = x.to(torch.bfloat16).float()
    if mode == "bf16":
        return hi
    return hi + (x - hi).to(torch.bfloat16).float()


def tanh_ap(x, on):
    """tanh.approx.f32 model: relative error up to 2^-11 (PTX ISA), emulated as truncation to 11 mantissa bits."""
    y = torch.tanh(x)
    if not on:
        return y
    return (y.view(torch.int32) & ~0xFFF).view(torch.float32)


def conv(x, W, bias, kf, kt, C, cfg, x2=None, C2=0, pool=False):
    N = W.shape[0]
    Wr = rnd(W, cfg.get("w_conv"))
    w = Wr[:, :kf * kt * C].reshape(N, kf, kt, C).permute(0, 3, 1, 2)
    y = F.conv2d(x.permute(0, 3, 2, 1), w, None, padding=(kf // 2, kt // 2))
    if x2 is not None:
        y = y + F.conv2d(x2.permute(0, 3, 2, 1), Wr[:, kf * kt * C:].reshape(N, C2, 1, 1), None)
    y = (y + bias.view(1, -1, 1, 1)).relu()
    if pool:
        y = F.max_pool2d(y, (2, 1))
    return y.permute(0, 3, 2, 1).contiguous()


def lstm(gx, whh, H, reverse, cfg):
    B, T, _ = gx.shape
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    out = torch.zeros(B, T, H)
    Wt = rnd(whh.float(), cfg.get("w_hh")).t().contiguous()
    ap = cfg.get("tanhap")
    for s in range(T):
        t = T - 1 - s if reverse else s
        gates = gx[:, t] + rnd(h, cfg.get("h_fb")) @ Wt
        g4 = gates.view(B, H // 32, 32, 4)
        i, f, g, o = [g4[..., k].reshape(B, H) for k in range(4)]
        if ap:      # the kernel: sigmoid(x) = 0.5 tanh.approx(0.5 x) + 0.5
            si, sf, so = [0.5 * tanh_ap(0.5 * v, True) + 0.5 for v in (i, f, o)]
        else:
            si, sf, so = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o)
        c = sf * c + si * tanh_ap(g, ap)
        h = so * tanh_ap(c, ap)
        out[:, t] = h
    return out


@torch.no_grad()
def forward(P, x, H, layers, cfg):
    """P: packed weights in FP32 (packing.pack_state_dict(..., weight_dtype=torch.float32)); returns frame logits + internals."""
    B, _, _, T = x.shape
    A = lambda t: rnd(t, cfg.get("a_conv"))
    y = F.conv2d(x, P["conv1.w"].view(32, 1, 3, 3), P["conv1.b"], padding=1).relu()
    act1 = A(F.max_pool2d(y, (2, 1)).permute(0, 3, 2, 1).contiguous())
    h1 = A(conv(act1, P["res1.c1.w"], P["res1.c1.b"], 3, 3, 32, cfg))
    act2 = A(conv(h1, P["res1.c2.w"], P["res1.c2.b"], 3, 3, 64, cfg, act1, 32, pool=True))
    h2 = A(conv(act2, P["res2.c1.w"], P["res2.c1.b"], 3, 3, 64, cfg))
    act3 = A(conv(h2, P["res2.c2.w"], P["res2.c2.b"], 3, 3, 128, cfg, act2, 64))
    feat = rnd(conv(act3, P["freq.w"], P["freq.b"], 7, 3, 128, cfg, pool=True), cfg.get("feat"))
    Hl = H // 2
    D = 2 * H + 2 * Hl
    xin = feat.reshape(B, T, -1)
    rnn = torch.zeros(B, T, D)
    for l in range(layers):
        gx = xin @ rnd(P[f"rnn{l}.wih"], cfg.get("w_ih")).t() + P[f"rnn{l}.b"]
        outs = [lstm(gx[..., d * 4 * H:(d + 1) * 4 * H], P[f"rnn{l}.whh{d}"], H, d, cfg) for d in range(2)]
        if l == 0:
            for d in range(2):
                o = 8 * H + d * 4 * Hl
                rnn[..., 2 * H + d * Hl:2 * H + (d + 1) * Hl] = lstm(gx[..., o:o + 4 * Hl], P[f"loc.whh{d}"], Hl, d, cfg)
        cat = torch.cat(outs, dim=-1)
        if l == layers - 1:
            rnn[..., :2 * H] = cat
        xin = rnd(cat, cfg.get("seq"))
    internals = {"feat": feat, "rnn": rnn}
    head_in = rnd(rnn, cfg.get("rnn_out"))
    qkv = rnd(head_in @ rnd(P["attn.qkv.w"], cfg.get("w_attn")).t() + P["attn.qkv.b"], cfg.get("qkv"))
    hd = D // 8
    q, k, v = qkv.reshape(B, T, 3, 8, hd).permute(2, 0, 3, 1, 4)
    s = torch.clamp((q @ k.transpose(-2, -1)) * hd ** -0.5, -10, 10)
    if cfg.get("p_att"):        # the kernel: P = exp(s - 10) as bf16 (unnormalised), row sums of the fp32 values, O / rowsum
        e = torch.exp(s - 10.0)
        o = (rnd(e, cfg.get("p_att")) @ v) / e.sum(-1, keepdim=True)
    else:
        o = s.softmax(-1) @ v
    att = rnd(o.transpose(1, 2).reshape(B, T, D), cfg.get("att"))
    proj = att @ rnd(P["attn.proj.w"], cfg.get("w_attn")).t() + P["attn.proj.b"]
    normed = F.layer_norm(rnn + proj, (D,), P["ln.w"], P["ln.b"], eps=1e-6)
    internals["attn_norm"] = normed
    head_in = rnd(normed, cfg.get("normed"))
    shared = rnd((head_in @ rnd(P["fc1.w"], cfg.get("w_head")).t() + P["fc1.b"]).relu(), cfg.get("shared"))
    lg = shared @ rnd(P["heads.w"], cfg.get("w_head")).t() + P["heads.b"]
    return lg[..., :88].transpose(1, 2), internals


def main():
    from music_transcription_b200 import synth
    from music_transcription_b200.packing import pack_state_dict
    from oracle import frontend as ofe, model as omodel
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, nargs="+", default=[0, 3])
    ap.add_argument("--gain", type=float, default=3 ** -0.5)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--frames", type=int, default=938)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    H, L = args.hidden, args.layers
    sd = synth.synth_state_dict("cnn_rnn_large", 320, H, L, seed=1, gain=args.gain)
    P = pack_state_dict(sd, "cnn_rnn_large", 320, H, L, weight_dtype=torch.float32)
    n = (args.frames - 1) * 512
    mel = torch.stack([torch.from_numpy(ofe.logmel(synth.piano_chord(k, n_samples=n))) for k in args.chunks])[:, None]
    ref, ri = omodel.large_forward(sd, mel, H, L, return_internals=True)
    pref = torch.sigmoid(ref)

    def report(name, cfg):
        out, it = forward(P, mel, H, L, cfg)
        dp = (torch.sigmoid(out) - pref).abs()
        flips = ((torch.sigmoid(out) > 0.5) != (pref > 0.5)).float().mean().item()
        d_rnn = (it["rnn"] - ri["rnn"]).abs().max().item()
        d_norm = (it["attn_norm"] - ri["attn_norm"]).abs().max().item()
        d_feat = (it["feat"].permute(0, 3, 2, 1) - ri["freq"]).abs().max().item()
        print(f"{name:34s} prob max {dp.max():.2e} mean {dp.mean():.2e} | logit max {(out - ref).abs().max():.2e} | flips {100 * flips:.3f}% "
              f"| feat {d_feat:.1e} rnn {d_rnn:.1e} ln {d_norm:.1e}", flush=True)

    print(f"# chunks {args.chunks} gain {args.gain:.4f} hidden {H} layers {L} T {mel.shape[-1]}; logits std {ref.std():.3f}, "
          f"probs in [{pref.min():.3f}, {pref.max():.3f}]")
    report("exact fp32 emulation (packing only)", {})
    for s in SITES:
        report("only " + s + " = bf16", {s: "bf16"})
    allbf = {s: "bf16" for s in SITES}
    report("ALL bf16 (= the fast kernels)", allbf)
    report("ALL bf16, exact tanh", dict(allbf, tanhap=None))
    weights = ("w_conv", "w_ih", "w_hh", "w_attn", "w_head")
    report("weights bf16, activations exact", {s: "bf16" for s in weights})
    report("activations bf16, weights exact", {s: "bf16" for s in SITES if s not in weights})
    report("weights split, activations bf16", dict(allbf, **{s: "split" for s in weights}))
    report("ALL split (hi+lo), tanh.approx", dict({s: "split" for s in SITES}, tanhap="bf16"))
    report("ALL split (hi+lo), exact tanh", dict({s: "split" for s in SITES}, tanhap=None))
    # candidates for a 'precise' mode that leaves the conv stack and the big GEMM in plain bf16
    lstm_precise = dict(allbf, w_hh="split", h_fb="split", seq="split", tanhap=None, rnn_out="split")
    report("LSTM precise (whh,h,seq split; tanh)", lstm_precise)
    report("  + w_ih split", dict(lstm_precise, w_ih="split"))
    report("  + w_ih, feat split", dict(lstm_precise, w_ih="split", feat="split"))
    report("  + w_ih, feat, conv split", dict(lstm_precise, w_ih="split", feat="split", w_conv="split", a_conv="split"))


if __name__ == "__main__":
    main()
