"""Pack a reference checkpoint (``TranscriptionModel.state_dict()``, SURVEY.md
Appendix B) into the tensors the sm_100a kernels consume.  One-time work per
``load_state_dict``; pure tensor plumbing (torch on whatever device the
parameters live on), no inference arithmetic.

Packed names / layouts (all contiguous; "bf16 [N, K]" means K-major rows):

  conv1.w  f32 [32, 9]            stem conv, BatchNorm folded, tap = kf*3 + kt
  conv1.b  f32 [32]
  <conv>.w bf16 [Cout, taps*Cin (+Cskip)]   BN folded; K index = (kf, kt, cin), cin = 32
                                  (stem output) or a multiple of 64; for residual
                                  blocks the 1x1 skip conv (BN folded) is appended
                                  along K and its bias summed into <conv>.b
  <conv>.b f32 [Cout]
  rnn{l}.wih bf16 [Ngates, K]     all sequences of layer l stacked along N
                                  (main fwd, main bwd[, local fwd, local bwd]);
                                  rows in "slice order": slice s (32 hidden units),
                                  unit u, gate g -> row s*128 + 4*u + g; layer-0
                                  columns permuted from the reference's c*F + f
                                  feature index to the kernels' f*C + c
  rnn{l}.b  f32 [Ngates]          b_ih + b_hh, same row order
  rnn{l}.whh{d}, loc.whh{d}  bf16 [4H, H]   rows in slice order, columns natural
  attn.qkv.w/b, attn.proj.w/b, ln.w/b, fc1.w/b   natural nn.Linear layout (bf16 W, f32 b)
  heads.w  bf16 [Npad, K]         frame|onset|offset (or the single fc) rows, zero
                                  padded to a multiple of 128;  heads.b f32 [Npad]

BatchNorm folding (eval): y = (conv(x) - mean) / sqrt(var + 1e-5) * gamma + beta
(reference models/cnn_rnn_model.py:31,36,84-91,198).

PRECISE mode (``precise=True``; amt_model_config.precision = 1): every tensor-core contraction runs on
split-bf16 operands, x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits), as THREE products
hi*hi + lo*hi + hi*lo accumulated in fp32 by the same kernels: the K axis is tripled.  Activations are stored
per channel group g as [hi(g) | lo(g) | hi(g)] and the matching weight columns as [Wh(g) | Wh(g) | Wl(g)]
(``split_k``); the group is the channel count of a pixel for conv inputs and for the layer-0 LSTM projection
(whose input is the conv stack's [T][F][C] tile), the whole K for the other linears.  The 32-channel stem output
is padded to a 128-channel group [hi | lo | hi | 0].  ``rnn*.whh`` stay plain bf16 (the recurrence contributes
< 1e-4 to the probabilities, tests/attribution.py).
"""
from __future__ import annotations

from typing import Dict

import torch

BN_EPS = 1e-5


def _fold_bn(sd, conv: str, bn: str):
    w = sd[conv + ".weight"].double()
    b = sd[conv + ".bias"].double()
    scale = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + BN_EPS)
    w = w * scale.view(-1, 1, 1, 1)
    b = (b - sd[bn + ".running_mean"].double()) * scale + sd[bn + ".bias"].double()
    return w, b


def _pack_conv_k(w: torch.Tensor) -> torch.Tensor:
    """[Co, Ci, kf, kt] -> [Co, kf*kt*Cip] with K index (kf, kt, ci); ci is kept when it is 32 (one
    SWIZZLE_64B block) and zero-padded to a multiple of 64 otherwise."""
    co, ci, kf, kt = w.shape
    ci64 = ci if ci == 32 else (ci + 63) // 64 * 64
    out = torch.zeros(co, kf, kt, ci64, dtype=w.dtype, device=w.device)
    out[..., :ci] = w.permute(0, 2, 3, 1)
    return out.reshape(co, kf * kt * ci64)


def slice_order(H: int) -> torch.Tensor:
    """perm[new_row] = reference gate row (g*H + u) for new_row = s*128 + 4*ul + g, u = 32*s + ul."""
    s = torch.arange(H // 32).view(-1, 1, 1)
    ul = torch.arange(32).view(1, -1, 1)
    g = torch.arange(4).view(1, 1, -1)
    return (g * H + 32 * s + ul).reshape(-1)


def _feat_perm_cols(w: torch.Tensor, C: int, F: int) -> torch.Tensor:
    """columns c*F + f  ->  f*C + c"""
    n = w.shape[0]
    return w.view(n, C, F).permute(0, 2, 1).reshape(n, F * C)


def split_k(w: torch.Tensor, group: int) -> torch.Tensor:
    """[N, K] fp32/fp64 -> bf16 [N, K/group * parts * group]: per group of ``group`` columns [Wh | Wh | Wl]
    (+ a zero block when group == 32, so that the group is 128 columns = two 64-column K blocks)."""
    n, k = w.shape
    assert k % group == 0, (k, group)
    v = w.float().reshape(n, k // group, group)
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    parts = [hi, hi, lo] + ([torch.zeros_like(hi)] if group == 32 else [])
    return torch.cat(parts, dim=-1).reshape(n, -1).contiguous()


def split_act(x: torch.Tensor, group: int) -> torch.Tensor:
    """The activation side of ``split_k`` (used by tests): [..., K] fp32 -> bf16 [..., parts*K] as [hi | lo | hi (| 0)] per group."""
    v = x.float().reshape(*x.shape[:-1], x.shape[-1] // group, group)
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    parts = [hi, lo, hi] + ([torch.zeros_like(hi)] if group == 32 else [])
    return torch.cat(parts, dim=-1).reshape(*x.shape[:-1], -1).contiguous()


def pack_state_dict(sd: Dict[str, torch.Tensor], model_type: str, n_mels: int, hidden_size: int, num_layers: int,
                    use_attention: bool = True, use_onset_offset_heads: bool = True,
                    device=None, weight_dtype=torch.bfloat16, precise: bool = False) -> Dict[str, torch.Tensor]:
    mt = model_type.lower()
    large = mt in ("cnn_rnn_large", "large")
    if not large and mt not in ("cnn_rnn", "cnn+rnn"):
        raise ValueError(f"Unknown model type: {model_type}")
    H = hidden_size
    out: Dict[str, torch.Tensor] = {}
    bf = weight_dtype          # bf16 for the kernels; tests/attribution.py packs fp32 to isolate rounding steps

    def put(name, t, dtype, group=None):
        """group: K-axis channel group(s) of a contraction weight -- an int, or [(columns, group), ...] for a weight
        made of several K segments (conv + appended skip conv); None for biases / recurrent weights."""
        if precise and group is not None:
            segs = [(t.shape[1], group)] if isinstance(group, int) else group
            cols, parts = 0, []
            for width, g in segs:
                parts.append(split_k(t[:, cols:cols + width], g))
                cols += width
            assert cols == t.shape[1]
            t = torch.cat(parts, dim=1)
        t = t.to(dtype).contiguous()
        out[name] = t.to(device) if device is not None else t

    stem_conv, stem_bn = ("model.conv1.0", "model.conv1.1") if large else ("model.cnn.0", "model.cnn.1")
    w, b = _fold_bn(sd, stem_conv, stem_bn)
    put("conv1.w", w.reshape(32, 9), torch.float32)
    put("conv1.b", b, torch.float32)

    if large:
        for name, blk in (("res1", "model.res_block1"), ("res2", "model.res_block2")):
            w1, b1 = _fold_bn(sd, blk + ".conv1", blk + ".bn1")
            cin = w1.shape[1]
            put(name + ".c1.w", _pack_conv_k(w1), bf, group=cin)
            put(name + ".c1.b", b1, torch.float32)
            w2, b2 = _fold_bn(sd, blk + ".conv2", blk + ".bn2")
            ws, bs = _fold_bn(sd, blk + ".skip.0", blk + ".skip.1")
            k2, ks = _pack_conv_k(w2), _pack_conv_k(ws)
            put(name + ".c2.w", torch.cat([k2, ks], dim=1), bf, group=[(k2.shape[1], w2.shape[1]), (ks.shape[1], cin)])
            put(name + ".c2.b", b2 + bs, torch.float32)
        w, b = _fold_bn(sd, "model.freq_aware_conv.0", "model.freq_aware_conv.1")
        put("freq.w", _pack_conv_k(w), bf, group=128)
        put("freq.b", b, torch.float32)
        C, F = 256, n_mels // 8
        rnn = "model.rnn_main"
    else:
        w, b = _fold_bn(sd, "model.cnn.4", "model.cnn.5")
        put("c2.w", _pack_conv_k(w), bf, group=32)
        put("c2.b", b, torch.float32)
        C, F = 64, n_mels // 4
        rnn = "model.rnn"

    perm = slice_order(H)
    for l in range(num_layers):
        ws, bs = [], []
        for d, suf in enumerate(("", "_reverse")):
            wih = sd[f"{rnn}.weight_ih_l{l}{suf}"].float()
            if l == 0:
                wih = _feat_perm_cols(wih, C, F)
            ws.append(wih[perm.to(wih.device)])
            bias = (sd[f"{rnn}.bias_ih_l{l}{suf}"].double() + sd[f"{rnn}.bias_hh_l{l}{suf}"].double())
            bs.append(bias[perm.to(bias.device)])
            put(f"rnn{l}.whh{d}", sd[f"{rnn}.weight_hh_l{l}{suf}"].float()[perm.to(wih.device)], bf)
        if large and l == 0:
            Hl = H // 2
            perm_l = slice_order(Hl)
            for d, suf in enumerate(("", "_reverse")):
                wih = _feat_perm_cols(sd[f"model.rnn_local.weight_ih_l0{suf}"].float(), C, F)
                ws.append(wih[perm_l.to(wih.device)])
                bias = (sd[f"model.rnn_local.bias_ih_l0{suf}"].double() + sd[f"model.rnn_local.bias_hh_l0{suf}"].double())
                bs.append(bias[perm_l.to(bias.device)])
                put(f"loc.whh{d}", sd[f"model.rnn_local.weight_hh_l0{suf}"].float()[perm_l.to(wih.device)], bf)
        put(f"rnn{l}.wih", torch.cat(ws, dim=0), bf, group=C if l == 0 else 2 * H)
        put(f"rnn{l}.b", torch.cat(bs, dim=0), torch.float32)

    def pad_rows(w, b, mult=128):
        n = w.shape[0]
        npad = (n + mult - 1) // mult * mult
        wp = torch.zeros(npad, w.shape[1], dtype=w.dtype, device=w.device)
        bp = torch.zeros(npad, dtype=b.dtype, device=b.device)
        wp[:n], bp[:n] = w, b
        return wp, bp

    if large:
        if use_attention:
            put("attn.qkv.w", sd["model.attention.qkv.weight"], bf, group=sd["model.attention.qkv.weight"].shape[1])
            put("attn.qkv.b", sd["model.attention.qkv.bias"], torch.float32)
            put("attn.proj.w", sd["model.attention.proj.weight"], bf, group=sd["model.attention.proj.weight"].shape[1])
            put("attn.proj.b", sd["model.attention.proj.bias"], torch.float32)
            put("ln.w", sd["model.attention_norm.weight"], torch.float32)
            put("ln.b", sd["model.attention_norm.bias"], torch.float32)
        if use_onset_offset_heads:
            put("fc1.w", sd["model.shared_fc.weight"], bf, group=sd["model.shared_fc.weight"].shape[1])
            put("fc1.b", sd["model.shared_fc.bias"], torch.float32)
            w = torch.cat([sd[f"model.{n}_head.weight"].float() for n in ("frame", "onset", "offset")], dim=0)
            b = torch.cat([sd[f"model.{n}_head.bias"].float() for n in ("frame", "onset", "offset")], dim=0)
        else:
            w, b = sd["model.fc.weight"].float(), sd["model.fc.bias"].float()
    else:
        w, b = sd["model.fc.weight"].float(), sd["model.fc.bias"].float()
    wp, bp = pad_rows(w, b)
    put("heads.w", wp, bf, group=wp.shape[1])
    put("heads.b", bp, torch.float32)
    return out
