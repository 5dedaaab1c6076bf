"""CPU: packed-weight layouts (music_transcription_b200/packing.py) checked by running a torch
emulation of the kernels' contracts (tests/emulate.py) against the reference golden outputs."""
import glob
import os

import numpy as np
import pytest
import torch

from music_transcription_b200 import synth
from music_transcription_b200.packing import pack_state_dict, slice_order
from tests.emulate import emu_forward

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_slice_order_is_a_permutation():
    for H in (64, 128, 256, 512):
        p = slice_order(H)
        assert sorted(p.tolist()) == list(range(4 * H))
        # new row s*128 + 4*u + g holds reference row g*H + 32*s + u
        assert p[1 * 128 + 4 * 5 + 2].item() == 2 * H + 32 * 1 + 5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*_*.npz"))),
                         ids=lambda p: os.path.basename(p)[6:-4])
def test_packed_emulation_matches_reference(path):
    g = np.load(path)
    n_mels, H, L, B, T, attn, heads, seed, xseed = [int(v) for v in g["cfg"]]
    mt = str(g["model_type"])
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=seed, use_attention=bool(attn), use_onset_offset_heads=bool(heads))
    P = pack_state_dict(sd, mt, n_mels, H, L, bool(attn), bool(heads))
    x = torch.from_numpy(g["x"])
    torch.set_num_threads(4)
    out = emu_forward(P, x, mt, n_mels, H, L, bool(attn), bool(heads), bf16_acts=True)
    for k, v in out.items():
        ref = torch.from_numpy(g[k])
        dl = (v - ref).abs().max().item()
        dp = (torch.sigmoid(v) - torch.sigmoid(ref)).abs()
        # bf16 weights + bf16 activations: same tolerance the GPU tests state
        assert dl < 0.2 and dp.max().item() < 4e-2 and dp.mean().item() < 4e-3, (k, dl, dp.max().item())


def test_fast_chord_batch_is_the_survey_recipe_and_pcm16_round_trip():
    ks = [0, 7, 13]
    fast = synth.piano_chord_batch_fast(ks, n_samples=20000)
    for i, k in enumerate(ks):
        assert np.array_equal(fast[i].numpy(), synth.piano_chord(k, n_samples=20000))
    pcm = synth.to_pcm16(fast)
    assert pcm.dtype == torch.int16 and (pcm.float() / 32768.0 - fast).abs().max() <= 0.5 / 32768 + 1e-9


def test_precise_packing_reconstructs_fp32_weights():
    """precise=True: every contraction weight is [Wh | Wh | Wl] per channel group; Wh + Wl must reproduce the fp32
    packed weight to 2^-16 relative, and the group structure must match split_act's [hi | lo | hi (| 0)]."""
    from music_transcription_b200.packing import split_act, split_k
    sd = synth.synth_state_dict("cnn_rnn_large", 64, 128, 2, seed=3)
    fast = pack_state_dict(sd, "cnn_rnn_large", 64, 128, 2, weight_dtype=torch.float32)
    prec = pack_state_dict(sd, "cnn_rnn_large", 64, 128, 2, precise=True)
    # x . w  ==  split_act(x) . split_k(w)  up to the dropped lo*lo term, for both group shapes
    g = torch.Generator().manual_seed(0)
    for group, k in ((32, 96), (64, 128), (256, 256)):
        w, x = torch.randn(5, k, generator=g), torch.randn(3, k, generator=g)
        got = split_act(x, group).double() @ split_k(w, group).double().t()
        assert (got - x.double() @ w.double().t()).abs().max() < 4 * k ** 0.5 * 2 ** -16      # bf16 alone: ~k^0.5 * 2^-8
    assert prec["rnn0.whh0"].shape == fast["rnn0.whh0"].shape            # recurrent weights stay plain bf16
    w = prec["attn.qkv.w"].float()
    K = fast["attn.qkv.w"].shape[1]
    assert w.shape[1] == 3 * K and torch.equal(w[:, :K], w[:, K:2 * K])
    assert ((w[:, :K] + w[:, 2 * K:]) - fast["attn.qkv.w"]).abs().max() < 2 ** -16
    assert prec["res1.c1.w"].shape[1] == 9 * 128 and prec["res1.c2.w"].shape[1] == 9 * 192 + 128
