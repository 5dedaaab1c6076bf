"""GPU parity tests of the individual sm_100a kernels, called through the C ABI
(ctypes) and compared with plain fp32 PyTorch math on the same (bf16-rounded) inputs
or with the oracle for integer work.  Run with: pytest -m gpu"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from music_transcription_b200 import _lib, synth
from music_transcription_b200.packing import slice_order

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _stream():
    return _lib.stream_ptr(torch.device(DEV))


def _bf(x):
    return x.to(torch.bfloat16)


# ----------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,relu,f32", [
    (128, 256, 64, 0, 1),
    (200, 128, 128, 1, 0),
    (1000, 64, 192, 0, 1),
    (938 * 2, 512, 1024, 1, 0),
    (300, 384, 1536, 0, 1),
    (4100, 1536, 640, 0, 0),
    (129, 192, 64, 0, 0),            # second CTA of the pair owns a single valid row
    (60032, 512, 256, 0, 0),         # config-[2]/[3] row count (64 x 938): 234.5 pair tiles
])
def test_gemm_tcgen05_matches_fp32_matmul(M, N, K, relu, f32):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = _bf(torch.randn(M, K, generator=g)).to(DEV)
    W = _bf(torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if f32 else torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().amt_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, N, K, N, relu, f32,
                                        _stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    if relu:
        ref = ref.relu()
    tol = 2e-3 if f32 else 2e-2
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < tol


def test_gemm_rejects_bad_shapes():
    a = torch.zeros(128, 96, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().amt_gemm_bf16(_lib.ptr(a), _lib.ptr(a), _lib.ptr(a), _lib.ptr(a), 128, 128, 96, 128, 0, 0,
                                            _stream()))


# ----------------------------------------------------------------------------- conv
def _conv_ref(x_btfc, w, bias, kf, kt, relu, pool, x2=None, w2=None):
    # x [B,T,F,C] -> NCHW with H=F, W=T
    x = x_btfc.float().permute(0, 3, 2, 1)
    y = F.conv2d(x, w.float(), None, padding=(kf // 2, kt // 2))
    if x2 is not None:
        y = y + F.conv2d(x2.float().permute(0, 3, 2, 1), w2.float(), None)
    y = y + bias.view(1, -1, 1, 1)
    if relu:
        y = y.relu()
    if pool:
        y = F.max_pool2d(y, (2, 1))
    return y.permute(0, 3, 2, 1).contiguous()        # [B,T,F',Co]


def _pack_w(w):     # [Co,Ci,kf,kt] -> [Co, kf*kt*Ci]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("B,T,Fq,Ci,Co,kf,kt,pool,skip", [
    (1, 16, 16, 64, 64, 3, 3, 0, 0),
    (2, 37, 20, 64, 128, 3, 3, 1, 0),
    (1, 50, 9, 128, 256, 7, 3, 1, 0),
    (2, 21, 40, 128, 128, 3, 3, 0, 64),
    (1, 938, 80, 64, 64, 3, 3, 1, 64),
    (2, 33, 24, 32, 64, 3, 3, 0, 0),          # 32-channel input (stem output): 64-byte rows, SWIZZLE_64B
    (1, 70, 40, 64, 64, 3, 3, 1, 32),         # 64-channel main + 32-channel skip source
    (1, 19, 160, 32, 64, 3, 3, 1, 0),
    (1, 40, 80, 128, 256, 7, 3, 1, 0),
])
def test_conv_implicit_gemm_matches_conv2d(B, T, Fq, Ci, Co, kf, kt, pool, skip):
    g = torch.Generator().manual_seed(B * 1000 + T + Fq + Ci + Co)
    x = _bf(torch.randn(B, T, Fq, Ci, generator=g)).to(DEV)
    w = _bf(torch.randn(Co, Ci, kf, kt, generator=g) / (Ci * kf * kt) ** 0.5).to(DEV)
    bias = torch.randn(Co, generator=g).to(DEV)
    x2 = w2 = None
    wp = _pack_w(w)
    if skip:
        x2 = _bf(torch.randn(B, T, Fq, skip, generator=g)).to(DEV)
        w2 = _bf(torch.randn(Co, skip, 1, 1, generator=g) / skip ** 0.5).to(DEV)
        wp = torch.cat([wp, _pack_w(w2)], dim=1).contiguous()
    Fo = Fq // 2 if pool else Fq
    out = torch.full((B, T, Fo, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().amt_conv_bf16(_lib.ptr(x), _lib.ptr(x2), _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(out), B, T, Fq,
                                        Ci, skip, Co, kf, kt, 1, pool, _stream()))
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, bias, kf, kt, True, pool, x2, w2)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("B,T,Fq,Ci,Co,kf,kt,pool,skip", [
    (2, 33, 24, 32, 64, 3, 3, 0, 0),          # stem output: 32 channels -> the 128-channel group [hi | lo | hi | 0]
    (1, 70, 40, 64, 64, 3, 3, 1, 32),         # main 64 -> 192 channels, skip 32 -> 128
    (2, 21, 40, 128, 128, 3, 3, 0, 64),       # main 128 -> 384, skip 64 -> 192
    (1, 50, 9, 128, 256, 7, 3, 1, 0),         # the 7x3 frequency-aware conv, pooled
])
def test_conv_split_bf16_precise_mode(B, T, Fq, Ci, Co, kf, kt, pool, skip):
    """Precise mode: fp32 inputs and weights enter as split-bf16 operands ([hi | lo | hi] activations,
    [Wh | Wh | Wl] weights: three MMA products), the epilogue emits [hi | lo | hi] again.  hi + lo must match the
    fp32 convolution of the UNROUNDED operands to ~2^-16 relative (the bf16 path is at 2^-8), and the two hi
    copies must be identical."""
    from music_transcription_b200.packing import split_act, split_k
    g = torch.Generator().manual_seed(B * 1000 + T + Fq + Ci + Co)
    x = torch.randn(B, T, Fq, Ci, generator=g)
    w = torch.randn(Co, Ci, kf, kt, generator=g) / (Ci * kf * kt) ** 0.5
    bias = torch.randn(Co, generator=g)
    xs = split_act(x, Ci).to(DEV)
    wp = split_k(_pack_w(w), Ci)
    Cs = xs.shape[-1]
    x2 = w2 = x2s = None
    C2s = 0
    if skip:
        x2 = torch.randn(B, T, Fq, skip, generator=g)
        w2 = torch.randn(Co, skip, 1, 1, generator=g) / skip ** 0.5
        x2s = split_act(x2, skip).to(DEV)
        C2s = x2s.shape[-1]
        wp = torch.cat([wp, split_k(_pack_w(w2), skip)], dim=1)
    wp = wp.contiguous().to(DEV)
    Fo = Fq // 2 if pool else Fq
    out = torch.full((B, T, Fo, 3 * Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().amt_conv_bf16(_lib.ptr(xs), _lib.ptr(x2s), _lib.ptr(wp), _lib.ptr(bias.to(DEV)), _lib.ptr(out), B, T, Fq,
                                        Cs, C2s, Co, kf, kt, 1, pool | 2, _stream()))
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, bias, kf, kt, True, pool, x2, w2)
    o = out.float().cpu()
    assert torch.isfinite(o).all()
    hi, lo, hi2 = o[..., :Co], o[..., Co:2 * Co], o[..., 2 * Co:]
    assert torch.equal(hi, hi2)
    assert ((hi + lo) - ref).abs().max().item() < 2e-4          # bf16 operands alone: ~2e-2 on these magnitudes
    assert (hi - ref).abs().max().item() > 10 * ((hi + lo) - ref).abs().max().item()   # the lo part really carries the residual


def test_split3_kernel_matches_packing_split_act():
    from music_transcription_b200.packing import split_act
    g = torch.Generator().manual_seed(2)
    x = torch.randn(77, 1536, generator=g) * 3
    out = torch.empty(77, 3 * 1536, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().amt_split3_bf16(_lib.ptr(x.to(DEV)), 1, _lib.ptr(out), 77, 1536, _stream()))
    assert torch.equal(out.cpu(), split_act(x, 1536))
    xb = x.to(torch.bfloat16)
    _lib.check(_lib.lib().amt_split3_bf16(_lib.ptr(xb.to(DEV)), 0, _lib.ptr(out), 77, 1536, _stream()))
    o = out.cpu()
    assert torch.equal(o[:, :1536], xb) and torch.equal(o[:, 3072:], xb) and (o[:, 1536:3072] == 0).all()


def test_conv_rejects_unsupported_shapes():
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(64, 9 * 64, dtype=torch.bfloat16, device=DEV)
    b = torch.zeros(64, device=DEV)
    y = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=DEV)
    L = _lib.lib()
    for bad in (dict(Cin=48), dict(Cout=96), dict(kf=5), dict(kt=1), dict(Cin=32, Cin2=64)):
        a = dict(Cin=64, Cin2=0, Cout=64, kf=3, kt=3)
        a.update(bad)
        with pytest.raises(ValueError):
            _lib.check(L.amt_conv_bf16(_lib.ptr(x), _lib.ptr(x) if a["Cin2"] else 0, _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), 1, 8, 8,
                                       a["Cin"], a["Cin2"], a["Cout"], a["kf"], a["kt"], 1, 0, _stream()))


# ----------------------------------------------------------------------------- LSTM recurrence
def _lstm_ref(gx, whh, reverse):
    """gx [B,T,4H] (gate order i,f,g,o, natural), whh [4H,H] bf16; h fed back in bf16 like the kernel."""
    B, T, G = gx.shape
    H = G // 4
    h = torch.zeros(B, H, device=gx.device)
    c = torch.zeros(B, H, device=gx.device)
    out = torch.zeros(B, T, H, device=gx.device)
    W = whh.float()
    for s in range(T):
        t = T - 1 - s if reverse else s
        gates = gx[:, t] + h.to(torch.bfloat16).float() @ W.t()
        i, f, g, o = gates.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


@pytest.mark.parametrize("B,T,Hs", [(3, 12, (128,)), (16, 40, (128, 128, 64, 64)), (20, 25, (256, 256)), (5, 30, (512, 512, 256, 256)),
                                    (40, 30, (512, 512)),                    # two batch groups of 32 chunks
                                    (100, 20, (512, 512, 256, 256)),         # more chunks than co-resident clusters at 32: BC = 64
                                    (3, 10, (640, 640))])                    # 20 slices > one cluster: cooperative L2 fallback
def test_lstm_recurrence_matches_stepwise_reference(B, T, Hs):
    g = torch.Generator().manual_seed(B * 100 + T)
    L = _lib.lib()
    n = len(Hs)
    ncols = sum(4 * h for h in Hs)
    gx_nat = [torch.randn(B, T, 4 * h, generator=g).to(DEV) for h in Hs]
    whh = [_bf(torch.randn(4 * h, h, generator=g) / h ** 0.5).to(DEV) for h in Hs]
    gx = torch.empty(B * T, ncols, device=DEV)
    wout = sum(Hs)
    out_bf = torch.full((B * T, wout), float("nan"), dtype=torch.bfloat16, device=DEV)
    out_f32 = torch.full((B * T, wout), float("nan"), device=DEV)
    seqs = (_lib.LstmSeq * n)()
    keep = []
    col = ocol = 0
    for i, h in enumerate(Hs):
        perm = slice_order(h).to(DEV)
        gx[:, col:col + 4 * h] = gx_nat[i].reshape(B * T, 4 * h)[:, perm]
        wp = whh[i][perm].contiguous()
        keep.append(wp)
        s = seqs[i]
        s.whh = _lib.ptr(wp)
        s.gx = gx.data_ptr() + col * 4
        s.out_bf16 = out_bf.data_ptr() + ocol * 2
        s.out_f32 = out_f32.data_ptr() + ocol * 4
        s.H, s.reverse, s.ld_gx, s.ld_out, s.ld_out32 = h, i % 2, ncols, wout, wout
        col += 4 * h
        ocol += h
    nbytes = L.amt_lstm_scratch_bytes(seqs, n, B)
    assert nbytes > 0
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.check(L.amt_lstm_recurrence(seqs, n, B, T, _lib.ptr(scratch), nbytes, _stream()))
    torch.cuda.synchronize()
    ocol = 0
    for i, h in enumerate(Hs):
        ref = _lstm_ref(gx_nat[i], whh[i], i % 2).reshape(B * T, h)
        got = out_f32[:, ocol:ocol + h]
        assert torch.isfinite(got).all()
        assert (got - ref).abs().max().item() < 5e-3, (i, h)
        assert (out_bf[:, ocol:ocol + h].float() - got).abs().max().item() < 8e-3
        ocol += h


# ----------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,T,hd", [(1, 64, 48), (2, 100, 48), (1, 938, 192), (2, 77, 96), (1, 130, 144),
                                     (3, 300, 192), (2, 129, 64), (1, 500, 128), (20, 938, 192)])
def test_attention_matches_clamped_softmax(B, T, hd):
    heads = 8
    D = heads * hd
    g = torch.Generator().manual_seed(T + hd)
    qkv = _bf(torch.randn(B * T, 3 * D, generator=g) * 1.5).to(DEV)
    out = torch.full((B * T, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().amt_attention_bf16(_lib.ptr(qkv), _lib.ptr(out), B, T, heads, hd, 10.0, _stream()))
    torch.cuda.synchronize()
    x = qkv.float().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    a = torch.clamp((q @ k.transpose(-2, -1)) * hd ** -0.5, -10.0, 10.0).softmax(-1)
    ref = (a @ v).transpose(1, 2).reshape(B * T, D)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 3e-2


# ----------------------------------------------------------------------------- notes / F1 (bit-exact)
def test_notes_kernel_is_bit_exact_with_oracle_and_reference_golden():
    from music_transcription_b200.pipeline import extract_notes
    from oracle import notes as onotes
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "notes_reference.npz"))
    fs = 16000 / 512
    for name in ("random", "sparse", "full", "empty", "edges", "seam"):
        roll = g[name + "_roll"].astype(np.float32)
        got = extract_notes(torch.from_numpy(roll).to(DEV), threshold=0.0)
        assert np.array_equal(got, onotes.group_notes(roll)), name
        assert np.array_equal(got[:, 0] + 21, g[name + "_pitch"])
        assert np.array_equal(got[:, 1] / fs, g[name + "_start"]) and np.array_equal(got[:, 2] / fs, g[name + "_end"])
    # seam merge through the segmented view (no concatenated copy on the device)
    segs = torch.from_numpy(np.stack([g["seam_a"], g["seam_b"]]).astype(np.float32)).to(DEV)
    assert np.array_equal(extract_notes(segs, 0.0), onotes.group_notes(g["seam_roll"].astype(np.float32)))
    # probabilities with values planted exactly on float32(threshold)
    for thr in (0.5, 0.1, 0.35000000000000003):
        p = synth.planted_probs(88, 938, [thr], seed=3, frac=0.05)
        got = extract_notes(torch.from_numpy(p).to(DEV), threshold=thr)
        assert np.array_equal(got, onotes.group_notes(onotes.threshold_roll(p, thr)))


def test_f1_counts_kernel_is_bit_exact():
    from music_transcription_b200 import evaluate as ev
    from oracle import f1 as of1
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "f1_reference.npz"))
    probs, rolls, lengths = g["probs"], g["rolls"].astype(np.float32), g["lengths"]
    thr = list(np.linspace(0.01, 0.99, 100)) + [0.35000000000000003, 0.5, 0.5]
    got = ev.f1_counts(torch.from_numpy(probs).to(DEV), torch.from_numpy(rolls).to(DEV), lengths, thr)
    want = of1.counts_grid(probs, rolls, lengths, thr)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    P, Y = torch.from_numpy(probs).to(DEV), torch.from_numpy(rolls).to(DEV)

    def mean_at(t):
        return float(np.mean(ev.f1_from_counts(ev.f1_counts(P, Y, lengths, [t])[:, 0])))
    for t, f in zip(g["at_t"], g["at_f1"]):
        assert mean_at(float(t)) == pytest.approx(float(f), abs=1e-15)
    best_t, best_f1 = ev.threshold_schedule_walk(mean_at)
    assert best_t == float(g["best_t"]) and best_f1 == pytest.approx(float(g["best_f1"]), abs=1e-15)


def test_f1_counts_config5_shape_properties():
    """50 pieces x 100 thresholds (BASELINE config 5 on one GPU): checksum identities."""
    from music_transcription_b200 import evaluate as ev
    n, T = 50, 938
    lens = np.array([937, 938, 469][0:1] * n, dtype=np.int32)
    lens[1::3], lens[2::3] = 938, 469
    thr = np.linspace(0.01, 0.99, 100)
    P = torch.stack([torch.from_numpy(synth.planted_probs(88, T, thr, seed=i)) for i in range(n)]).to(DEV)
    Y = torch.stack([torch.from_numpy(synth.bernoulli_roll(88, T, 0.05, seed=i)) for i in range(n)]).to(DEV)
    c = ev.f1_counts(P, Y, lens, thr)
    assert c.shape == (n, 100, 3)
    pos = np.array([int(Y[i, :, :lens[i]].sum().item()) for i in range(n)])
    assert np.array_equal(c[:, :, 0] + c[:, :, 2], np.repeat(pos[:, None], 100, 1))     # TP + FN = positives
    assert np.all(np.diff(c[:, :, 0] + c[:, :, 1], axis=1) <= 0)                           # predictions shrink with t
    from oracle import f1 as of1
    for i in (0, 7, 49):
        assert np.array_equal(c[i], of1.counts_grid([P[i].cpu().numpy()], [Y[i].cpu().numpy()], [lens[i]], thr)[0])


def test_sigmoid_threshold_strict_compare():
    x = torch.tensor([0.0, -2.1972246, 5.0, -5.0], device=DEV)
    probs, roll = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.lib().amt_sigmoid_threshold(_lib.ptr(x), 4, 0.5, _lib.ptr(probs), _lib.ptr(roll), _stream()))
    assert roll.tolist() == [0.0, 0.0, 1.0, 0.0]          # sigmoid(0) == 0.5 is NOT > 0.5
    assert torch.allclose(probs, torch.sigmoid(x), atol=1e-7)


def test_pack_roll_bits_equal_float_roll():
    """amt_pack_roll_u32: bit t%32 of word t/32 == the float roll of amt_sigmoid_threshold, for probabilities and for
    logits (same sigmoid), ragged T (not a multiple of 32), values planted exactly on float32(threshold)."""
    from music_transcription_b200 import pipeline
    for T, thr in ((938, 0.5), (33, 0.35000000000000003), (1, 0.5), (64, 0.1)):
        p = torch.from_numpy(synth.planted_probs(88 * 3, T, [thr], seed=T, frac=0.05)).to(DEV).view(3, 88, T)
        bits = pipeline.pack_roll(p, thr)
        assert bits.shape == (3, 88, (T + 31) // 32) and bits.dtype == torch.int32
        want = (p > float(np.float32(thr))).float().cpu().numpy()
        assert np.array_equal(pipeline.unpack_roll(bits, T), want)
        logits = torch.randn(2, 88, T, device=DEV) * 3
        logits[0, 0, 0] = 0.0                                 # sigmoid(0) == 0.5 is not > 0.5
        roll = torch.empty_like(logits)
        _lib.check(_lib.lib().amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.5, 0, _lib.ptr(roll), _stream()))
        assert np.array_equal(pipeline.unpack_roll(pipeline.pack_roll(logits, 0.5, apply_sigmoid=True), T), roll.cpu().numpy())


def test_bits_notes_equals_float_notes_incl_seams():
    """amt_bits_notes on the bit-packed roll == amt_threshold_notes on the probabilities it was packed from, for a
    multi-segment roll with notes crossing the seams and T not a multiple of 32."""
    from music_transcription_b200 import pipeline
    from oracle import notes as onotes
    for n_seg, T in ((5, 938), (3, 33), (1, 64)):
        p = torch.from_numpy(np.stack([synth.planted_probs(88, T, [0.5], seed=10 * n_seg + i, frac=0.05) for i in range(n_seg)])).to(DEV)
        p[:, 10, :] = 0.9                                         # sounds through every seam
        if n_seg > 1:
            p[0, 20, T - 2:] = 0.9
            p[1, 20, :3] = 0.9
        want = pipeline.extract_notes(p, 0.5)
        got = pipeline.extract_notes_from_bits(pipeline.pack_roll(p, 0.5), T)
        assert np.array_equal(got, want)
        roll = np.concatenate([onotes.threshold_roll(x, 0.5) for x in p.cpu().numpy()], axis=1)
        assert np.array_equal(got, onotes.group_notes(roll))


def _runs_roll(rng, P, N, p_on, p_off):
    """A (P, N) {0,1} roll of random runs (two-state Markov chain): realistic note-like structure, incl. long notes."""
    r = np.zeros((P, N), np.float32)
    state = rng.random(P) < 0.3
    for t in range(N):
        flip = rng.random(P)
        state = np.where(state, flip >= p_off, flip < p_on)
        r[:, t] = state
    return r


@pytest.mark.parametrize("n_seg,T,with_offset", [(1, 938, True), (4, 938, True), (3, 70, False), (2, 1100, True), (5, 33, True), (1, 1, False)])
def test_onset_aware_notes_are_bit_exact_with_the_oracle(n_seg, T, with_offset):
    """amt_onset_notes (decoding of the frame + onset + offset heads) == oracle.notes.group_notes_onset_aware on
    the concatenated rolls: notes crossing segment seams and 1024-frame block seams (T = 1100), T not a multiple of 32, dense
    and sparse onsets, re-strikes inside a sounding note, frames no onset opened, notes still open at the end."""
    from music_transcription_b200 import pipeline
    from oracle import notes as onotes
    rng = np.random.default_rng(1000 * n_seg + T)
    P, N = 88, n_seg * T
    frame = _runs_roll(rng, P, N, 0.02, 0.03)
    onset = _runs_roll(rng, P, N, 0.03, 0.6)
    offset = _runs_roll(rng, P, N, 0.01, 0.7) if with_offset else None
    frame[5, :] = 1.0                                   # sounds through every seam; one onset at the very start
    onset[5, :] = 0.0
    onset[5, 0] = 1.0
    if N > 40:
        onset[6, :] = 0.0
        frame[6, :] = 1.0
        onset[6, N - 1] = 1.0                           # a note that starts on the last frame
        onset[7, :] = 1.0                               # onset head stuck high: ONE rising edge
    want = onotes.group_notes_onset_aware(frame, onset, offset)
    seg = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a.reshape(P, n_seg, T).transpose(1, 0, 2))).to(DEV)
    got = pipeline.extract_notes_onset_aware(seg(frame), seg(onset), seg(offset), 0.5, 0.5, 0.5, logits=False)
    assert got.dtype == np.int32 and np.array_equal(got, want), (got.shape, want.shape)
    if n_seg == 4:                                      # logits in, thresholds through the sigmoid, and a too-small cap
        lg = lambda a: (seg(a) * 8.0 - 4.0)
        assert np.array_equal(pipeline.extract_notes_onset_aware(lg(frame), lg(onset), lg(offset)), want)
        with pytest.raises(_lib.AmtError):
            pipeline.extract_notes_onset_aware(seg(frame), seg(onset), seg(offset), logits=False, cap=len(want) - 1)


def test_async_roll_gather_single_process():
    from music_transcription_b200 import pipeline, sharding
    T, n = 70, 5
    p = torch.rand(n, 88, T, device=DEV)
    p[:, 3, :] = 0.9
    want = pipeline.extract_notes(p, 0.5)
    g = sharding.AsyncRollGather(n, n, T, DEV)
    bits = pipeline.pack_roll(p, 0.5).cpu()
    tickets = [g.submit(bits), g.submit(bits)]
    with pytest.raises(RuntimeError):
        g.submit(bits)                                            # both slots in flight
    for t in tickets:
        assert np.array_equal(g.result(t), want)
    assert np.array_equal(g.result(g.submit(bits)), want)        # slots are reusable


def test_threshold_notes_takes_caller_scratch_and_rejects_a_short_one():
    L = _lib.lib()
    assert L.amt_threshold_notes_scratch_ints(3, 88) == 2 * 3 * 88
    p = torch.rand(3, 88, 50, device=DEV)
    notes = torch.empty(88 * 75, 3, dtype=torch.int32, device=DEV)
    counts = torch.empty(89, dtype=torch.int32, device=DEV)
    scratch = torch.empty(2 * 3 * 88 - 1, dtype=torch.int32, device=DEV)
    st = L.amt_threshold_notes(_lib.ptr(p), 3, 88, 50, p.stride(0), p.stride(1), 0.5, _lib.ptr(notes), notes.shape[0],
                               _lib.ptr(counts), _lib.ptr(scratch), scratch.numel(), _stream())
    assert st == _lib.AMT_ERR_WORKSPACE and b"scratch" in L.amt_last_error()


# ----------------------------------------------------------------------------- memory safety / determinism
# compute-sanitizer is not available on this GPU pool (it answers "compute-sanitizer is closed on this pool"), so the
# kernels' memory behaviour is checked directly: every output (and scratch) buffer sits between guard regions filled with a
# sentinel, outputs are pre-filled with a second sentinel, each kernel runs twice, and the test asserts that (i) no guard
# byte changed (no out-of-bounds write), (ii) every output element was written, (iii) both runs agree bit for bit (the
# cross-CTA mbarrier / DSMEM protocols of the GEMM, conv, LSTM-cluster and attention kernels have no data race that
# changes a result).  Shapes straddle tile boundaries (rows % 128 != 0, T % 64 != 0, odd F).
GUARD = 4096


class _Guarded:
    def __init__(self, nbytes):
        self.raw = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
        self.body = self.raw[GUARD:GUARD + nbytes]
        self.body.fill_(0xFF)                                    # bf16 / f32 NaN patterns: an unwritten element stays NaN

    def view(self, dtype, *shape):
        return self.body.view(dtype).view(*shape)

    def intact(self):
        return bool((self.raw[:GUARD] == 0xA5).all() and (self.raw[GUARD + self.body.numel():] == 0xA5).all())


def _twice(run, outs):
    res = []
    for _ in range(2):
        for o in outs:
            o.body.fill_(0xFF)
        run()
        torch.cuda.synchronize()
        res.append([o.body.clone() for o in outs])
    for o, a, b in zip(outs, res[0], res[1]):
        assert o.intact(), "guard region overwritten"
        assert torch.equal(a, b), "two runs differ"
    return res[0]


def test_guarded_buffers_gemm_conv_attention():
    L = _lib.lib()
    g = torch.Generator().manual_seed(0)
    M, N, K = 333, 192, 320
    a, w, b = _bf(torch.randn(M, K, generator=g)).to(DEV), _bf(torch.randn(N, K, generator=g)).to(DEV), torch.randn(N, generator=g).to(DEV)
    for f32 in (0, 1):
        out = _Guarded(M * N * (4 if f32 else 2))
        _twice(lambda: _lib.check(L.amt_gemm_bf16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), out.body.data_ptr(), M, N, K, N, 0, f32, _stream())), [out])
        o = out.view(torch.float32 if f32 else torch.bfloat16, M, N).float()
        assert torch.isfinite(o).all() and (o - (a.float() @ w.float().t() + b)).abs().max() < 0.5
    for B, T, Fq, Ci, Co, kf, pool, split in ((2, 37, 21, 64, 128, 3, 1, 0), (1, 19, 11, 128, 256, 7, 0, 0), (1, 23, 10, 32, 64, 3, 1, 1)):
        from music_transcription_b200.packing import split_act, split_k
        x, wt = torch.randn(B, T, Fq, Ci, generator=g), torch.randn(Co, kf * 3 * Ci, generator=g) / (Ci * kf * 3) ** 0.5
        xs, ws = (split_act(x, Ci), split_k(wt, Ci)) if split else (_bf(x), _bf(wt))
        xs, ws, bias = xs.to(DEV).contiguous(), ws.to(DEV).contiguous(), torch.randn(Co, generator=g).to(DEV)
        Fo = Fq // 2 if pool else Fq
        out = _Guarded(B * T * Fo * Co * (3 if split else 1) * 2)
        _twice(lambda: _lib.check(L.amt_conv_bf16(_lib.ptr(xs), 0, _lib.ptr(ws), _lib.ptr(bias), out.body.data_ptr(), B, T, Fq, xs.shape[-1], 0,
                                                  Co, kf, 3, 1, pool | (2 if split else 0), _stream())), [out])
        assert torch.isfinite(out.view(torch.bfloat16, -1).float()).all()
    for B, T, hd in ((2, 150, 192), (1, 70, 64), (1, 70, 48)):
        D = 8 * hd
        qkv = _bf(torch.randn(B * T, 3 * D, generator=g)).to(DEV)
        out = _Guarded(B * T * D * 2)
        _twice(lambda: _lib.check(L.amt_attention_bf16(_lib.ptr(qkv), out.body.data_ptr(), B, T, 8, hd, 10.0, _stream())), [out])
        assert torch.isfinite(out.view(torch.bfloat16, -1).float()).all()


def test_guarded_buffers_lstm_cluster_and_fallback():
    L = _lib.lib()
    g = torch.Generator().manual_seed(1)
    for B, T, Hs in ((5, 9, (512, 512, 256, 256)), (37, 7, (128, 128)), (3, 5, (640, 640))):
        seqs = (_lib.LstmSeq * len(Hs))()
        keep, outs = [], []
        for i, H in enumerate(Hs):
            whh = _bf(torch.randn(4 * H, H, generator=g) / H ** 0.5)[slice_order(H)].contiguous().to(DEV)
            gx = torch.randn(B * T, 4 * H, generator=g).to(DEV)
            ob, of = _Guarded(B * T * H * 2), _Guarded(B * T * H * 4)
            keep += [whh, gx]
            outs += [ob, of]
            seqs[i] = _lib.LstmSeq(_lib.ptr(whh), _lib.ptr(gx), ob.body.data_ptr(), of.body.data_ptr(), H, i & 1, 4 * H, H, H)
        nb = L.amt_lstm_scratch_bytes(seqs, len(Hs), B)
        scratch = _Guarded(nb)
        scratch.body.zero_()
        res = _twice(lambda: _lib.check(L.amt_lstm_recurrence(seqs, len(Hs), B, T, scratch.body.data_ptr(), nb, _stream())), outs)
        assert scratch.intact()
        for r in res[1::2]:
            assert torch.isfinite(r.view(torch.float32)).all() and r.view(torch.float32).abs().max() <= 1.0


def test_guarded_buffers_frontend_notes_pack():
    from music_transcription_b200 import pipeline
    L = _lib.lib()
    fe = pipeline.Frontend.get(device=DEV)
    n = 3 * 512 + 77                                             # T = 4 frames, ragged tail
    wav = torch.from_numpy(synth.piano_chord_batch([0, 1], n_samples=n)).to(DEV)
    out, cmax = _Guarded(2 * 320 * 4 * 4), _Guarded(2 * 4)
    _twice(lambda: _lib.check(L.amt_logmel_f32(fe._h, _lib.ptr(wav), 2, n, n, out.body.data_ptr(), 80.0, cmax.body.data_ptr(), _stream())), [out, cmax])
    assert torch.isfinite(out.view(torch.float32, -1)).all()
    T, n_seg = 70, 3
    p = torch.rand(n_seg, 88, T, device=DEV)
    cap = 88 * ((n_seg * T + 1) // 2)
    notes, counts, scratch, bits = _Guarded(cap * 12), _Guarded(89 * 4), _Guarded(2 * 88 * n_seg * 4), _Guarded(n_seg * 88 * 3 * 4)

    def run():
        _lib.check(L.amt_threshold_notes(_lib.ptr(p), n_seg, 88, T, p.stride(0), p.stride(1), 0.5, notes.body.data_ptr(), cap,
                                         counts.body.data_ptr(), scratch.body.data_ptr(), 2 * 88 * n_seg, _stream()))
        _lib.check(L.amt_pack_roll_u32(_lib.ptr(p), n_seg * 88, T, 0.5, 0, bits.body.data_ptr(), _stream()))
    res = _twice(run, [counts, bits])
    assert notes.intact() and scratch.intact()
    total = int(res[0].view(torch.int32)[88])
    assert 0 < total <= cap
