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
