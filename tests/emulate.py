"""Test helper (CPU): a torch emulation of what each CUDA kernel computes FROM THE PACKED
WEIGHTS, in the kernels' own data layouts ([B,T,F,C] activations, (tap, cin) K order, slice-ordered
gate rows, bf16 operands with fp32 accumulation).  It lets the CPU suite check the packing logic
(BN folding, permutations, padding) against the oracle without a GPU, and predicts the bf16 error
the GPU tests should see.  Never used by the product path."""
import torch
import torch.nn.functional as F


def _r(x, on):          # bf16 rounding of an activation that the kernels store as bf16
    return x.to(torch.bfloat16).float() if on else x


def emu_conv(x, W, bias, kf, kt, C, x2=None, C2=0, pool=False):
    """x [B,T,F,C] ; W [N, kf*kt*C + C2] with K index (kf, kt, c)."""
    N = W.shape[0]
    w = W[:, :kf * kt * C].float().reshape(N, kf, kt, C).permute(0, 3, 1, 2)
    y = F.conv2d(x.permute(0, 3, 2, 1), w, None, padding=(kf // 2, kt // 2))
    if x2 is not None:
        w2 = W[:, kf * kt * C:].float().reshape(N, C2, 1, 1)
        y = y + F.conv2d(x2.permute(0, 3, 2, 1), w2, None)
    y = (y + bias.view(1, -1, 1, 1)).relu()
    if pool:
        y = F.max_pool2d(y, (2, 1))
    return y.permute(0, 3, 2, 1).contiguous()


def emu_lstm(gx, whh, H, reverse, round_h=True):
    """gx [B,T,4H] columns in slice order; whh [4H,H] rows in slice order."""
    B, T, _ = gx.shape
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    out = torch.zeros(B, T, H)
    W = whh.float()
    for s in range(T):
        t = T - 1 - s if reverse else s
        gates = gx[:, t] + _r(h, round_h) @ W.t()
        g4 = gates.view(B, H // 32, 32, 4)
        i, f, g, o = [g4[..., k].reshape(B, H) for k in range(4)]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


@torch.no_grad()
def emu_forward(P, x, model_type, n_mels, H, layers, use_attention=True, use_heads=True, bf16_acts=True):
    """P: packed dict (CPU tensors).  x (B,1,n_mels,T) -> dict of (B,88,T) logits."""
    large = model_type.lower() in ("cnn_rnn_large", "large")
    B, _, _, T = x.shape
    R = lambda t: _r(t, bf16_acts)
    # stem conv as conv1_mma_kernel computes it (fast mode): x = hi + lo, w = wh + wl in bf16, the three products
    # hi*wh + lo*wh + hi*wl accumulated in fp32 (the fp32 stencil of the precise mode agrees with it to 2^-16)
    w1 = P["conv1.w"].view(32, 1, 3, 3).float()
    if bf16_acts:
        sp = lambda t: (t.to(torch.bfloat16).float(), (t - t.to(torch.bfloat16).float()).to(torch.bfloat16).float())
        (xh, xl), (wh, wl) = sp(x.float()), sp(w1)
        y = (F.conv2d(xh.double(), wh.double(), padding=1) + F.conv2d(xl.double(), wh.double(), padding=1)
             + F.conv2d(xh.double(), wl.double(), padding=1)).float() + P["conv1.b"].view(1, 32, 1, 1)
        y = y.relu()
    else:
        y = F.conv2d(x, w1, P["conv1.b"], padding=1).relu()
    y = F.max_pool2d(y, (2, 1)).permute(0, 3, 2, 1)                       # [B,T,F1,32]
    act1 = R(y.contiguous())                                              # 32 channels
    if large:
        h1 = R(emu_conv(act1, P["res1.c1.w"], P["res1.c1.b"], 3, 3, 32))
        act2 = R(emu_conv(h1, P["res1.c2.w"], P["res1.c2.b"], 3, 3, 64, act1, 32, pool=True))
        h2 = R(emu_conv(act2, P["res2.c1.w"], P["res2.c1.b"], 3, 3, 64))
        act3 = R(emu_conv(h2, P["res2.c2.w"], P["res2.c2.b"], 3, 3, 128, act2, 64))
        feat = R(emu_conv(act3, P["freq.w"], P["freq.b"], 7, 3, 128, pool=True))
        Hl = H // 2
    else:
        feat = R(emu_conv(act1, P["c2.w"], P["c2.b"], 3, 3, 32, pool=True))
        Hl = 0
    xin = feat.reshape(B, T, -1)
    D = 2 * H + 2 * Hl
    rnn = torch.zeros(B, T, D)
    for l in range(layers):
        gx = xin @ P[f"rnn{l}.wih"].float().t() + P[f"rnn{l}.b"]
        last = l == layers - 1
        outs = [emu_lstm(gx[..., d * 4 * H:(d + 1) * 4 * H], P[f"rnn{l}.whh{d}"], H, d, bf16_acts) for d in range(2)]
        if large and l == 0:
            for d in range(2):
                o = 8 * H + d * 4 * Hl
                rnn[..., 2 * H + d * Hl:2 * H + (d + 1) * Hl] = emu_lstm(gx[..., o:o + 4 * Hl], P[f"loc.whh{d}"], Hl, d, bf16_acts)
        cat = torch.cat(outs, dim=-1)
        if last:
            rnn[..., :2 * H] = cat
        xin = R(cat)
    head_in = R(rnn)
    if large and use_attention:
        qkv = R(head_in @ P["attn.qkv.w"].float().t() + P["attn.qkv.b"])
        hd = D // 8
        q, k, v = qkv.reshape(B, T, 3, 8, hd).permute(2, 0, 3, 1, 4)
        a = torch.clamp((q @ k.transpose(-2, -1)) * hd ** -0.5, -10, 10).softmax(-1)
        att = R((a @ v).transpose(1, 2).reshape(B, T, D))
        proj = att @ P["attn.proj.w"].float().t() + P["attn.proj.b"]
        head_in = R(F.layer_norm(rnn + proj, (D,), P["ln.w"], P["ln.b"], eps=1e-6))
    if large and use_heads:
        shared = R((head_in @ P["fc1.w"].float().t() + P["fc1.b"]).relu())
        lg = shared @ P["heads.w"].float().t() + P["heads.b"]
        return {n: lg[..., i * 88:(i + 1) * 88].transpose(1, 2) for i, n in enumerate(("frame", "onset", "offset"))}
    lg = head_in @ P["heads.w"].float().t() + P["heads.b"]
    return {"frame": lg[..., :88].transpose(1, 2)}
