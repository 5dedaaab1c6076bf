"""GPU parity of the drop-in boundary: log-mel frontend vs the oracle restatement,
TranscriptionModel (CUDA kernels through the C ABI) vs the oracle fp32 forward and vs the
golden outputs of the real reference, and the batched audio->notes pipeline.

Stated tolerances (measured values + ~20 %, profiles/r2_parity.md):
    log-mel  : max-abs <= 3e-3 dB, mean-abs <= 1e-4 dB vs the fp64-FFT oracle (fp32 FFT on the GPU)
    canonical config (n_mels 320 / hidden 512 / 3 layers, default-init-scale weights), vs the REAL reference's outputs:
               fast mode (bf16 operands, fp32 accumulation)  probabilities max-abs <= 2e-3, mean <= 3.5e-4, <= 0.4 % cells flip
               precise mode (split-bf16 operands)            probabilities max-abs <= 4e-4, mean <= 8e-5,   <= 0.1 % cells flip
               (tests/attribution.py: the fast-mode distance is the bf16 operand floor, spread evenly over the conv
               stack, the input projections and the heads; the recurrence contributes < 1e-4)
    logits   : max-abs <= 0.15, probabilities max-abs <= 3e-2, mean-abs <= 3e-3 on the stress
               checkpoints of synth.synth_state_dict (unit-gain weights, perturbed BatchNorm: logits
               reach +-3); a CPU emulation of the same bf16 roundings (tests/emulate.py) shows this is
               the bf16 operand floor (0.13 / 2.4e-2 / 2.5e-3), not kernel error -- the kernels
               themselves must agree with that emulation to 5e-2 / 1e-2 / 1e-3
    notes / TP-FP-FN counts: bit-exact given the same probability roll (test_gpu_kernels.py)
"""
import glob
import os

import numpy as np
import pytest
import torch

from music_transcription_b200 import pipeline, synth
from music_transcription_b200.transcription_model import TranscriptionModel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

LOGMEL_MAX_DB, LOGMEL_MEAN_DB = 3e-3, 1e-4          # dB, vs the fp64-FFT oracle (measured: <= 1.5e-3 max, <= 2.5e-5 mean)
LOGIT_MAX, PROB_MAX, PROB_MEAN = 0.15, 3e-2, 3e-3      # stress checkpoints; measured 0.124 / 2.36e-2 / 2.46e-3 (profiles/r2_parity.md)


def test_frontend_filterbank_matches_oracle():
    from oracle import frontend as ofe
    fe = pipeline.Frontend.get(device=DEV)
    fb = fe.filterbank()
    ref = ofe.mel_filterbank()
    assert fb.shape == ref.shape and np.abs(fb - ref).max() < 1e-7
    assert np.array_equal(fb > 0, ref > 0)


@pytest.mark.parametrize("case", ["chord0", "chord5", "silence", "noise", "short", "impulse"])
def test_logmel_matches_oracle(case):
    from oracle import frontend as ofe
    n = 480000
    if case.startswith("chord"):
        y = synth.piano_chord(int(case[5:]), n_samples=96000)
    elif case == "silence":
        y = np.zeros(32000, np.float32)
    elif case == "noise":
        y = np.random.default_rng(1).uniform(-1, 1, 48000).astype(np.float32)
    elif case == "short":
        y = synth.piano_chord(2, n_samples=3000)
    else:
        y = np.zeros(20000, np.float32)
        y[7777] = 1.0
    mel = pipeline.audio_to_mel(y, device=DEV)
    ref = ofe.logmel(y)
    assert mel.shape == (1, 1, 320, 1 + len(y) // 512) and mel.dtype == torch.float32
    d = np.abs(mel[0, 0].cpu().numpy() - ref)
    assert d.max() < LOGMEL_MAX_DB and d.mean() < LOGMEL_MEAN_DB, (case, d.max(), d.mean())


@pytest.mark.parametrize("n_mels", [229, 128, 64, 37, 512])
def test_logmel_other_filterbank_sizes(n_mels):
    """The reference class default is n_mels = 229 and checkpoints exist for other sizes: the ELL filterbank (rounds of 32
    filters, the round as wide as its widest band) must hold for few wide filters (37: up to ~90 bins per band), a
    partial last round (229, 37) and more filters than bins can separate (512: empty low filters)."""
    from oracle import frontend as ofe
    y = synth.piano_chord(7, n_samples=40000)
    mel = pipeline.audio_to_mel(y, n_mels=n_mels, device=DEV)
    ref = ofe.logmel(y, n_mels=n_mels)
    assert mel.shape == (1, 1, n_mels, 1 + len(y) // 512)
    d = np.abs(mel[0, 0].cpu().numpy() - ref)
    assert d.max() < LOGMEL_MAX_DB and d.mean() < LOGMEL_MEAN_DB, (n_mels, d.max(), d.mean())


def test_logmel_full_chunk_batch_shapes_and_floor():
    from oracle import frontend as ofe
    wav = torch.from_numpy(synth.piano_chord_batch([0, 1, 2])).to(DEV)
    fe = pipeline.Frontend.get(device=DEV)
    mel = fe.logmel(wav)
    assert mel.shape == (3, 1, 320, 938)
    for i in range(3):
        m = mel[i, 0]
        assert float(m.min()) >= float(m.max()) - 80.0 - 1e-4         # per-chunk top_db floor (main.py:125)
    ref = ofe.logmel(synth.piano_chord(1))
    d = np.abs(mel[1, 0].cpu().numpy() - ref)
    assert d.max() < LOGMEL_MAX_DB and d.mean() < LOGMEL_MEAN_DB, (d.max(), d.mean())


def test_deferred_floor_is_bitwise_the_floored_path():
    """Frontend.logmel(defer_floor=True) + amt_model_forward_db (the floor max(x, chunk max - 80 dB) applied by the stem
    convolution's load) must give bit for bit what flooring first gives -- incl. chunks where the floor bites hard (half of
    the chunk digital silence: -100 dB cells against a +20 dB maximum) and n_mels / batch shapes off the tile sizes."""
    fe = pipeline.Frontend.get(device=DEV)
    wav = torch.from_numpy(synth.piano_chord_batch([3, 4, 5], n_samples=64000)).to(DEV)
    wav[1, 20000:] = 0.0                               # floor active on most of this chunk
    wav[2] *= 1e-4                                     # a quiet chunk: its own maximum sets its floor
    mel = fe.logmel(wav)
    d = fe.logmel(wav, defer_floor=True)
    assert isinstance(d, pipeline.DeferredLogMel) and d.shape == mel.shape
    assert torch.equal(d.floored(), mel)
    assert float((d.mel < mel).float().mean()) > 0.05  # the deferred tensor really is unfloored
    for mt in ("cnn_rnn", "cnn_rnn_large"):
        m = TranscriptionModel(model_type=mt, n_mels=320, hidden_size=128, num_layers=1, device=DEV).eval()
        m.load_state_dict(synth.synth_state_dict(mt, 320, 128, 1, seed=2))
        assert torch.equal(m(d), m(mel)), mt
    with pytest.raises(ValueError):
        m(pipeline.DeferredLogMel(d.mel, d.chunk_max[:2], 80.0))


@pytest.mark.parametrize("n_mels,T,B", [(320, 938, 2), (37, 70, 3), (64, 129, 1)])
def test_stem_conv_tensor_path_matches_fp32_conv(n_mels, T, B):
    """conv1 (Conv 1->32 3x3 + folded BN + ReLU + MaxPool(2,1), reference cnn_rnn_model.py:179-182) runs on the tensor
    pipe in fast mode as a K = 27 split-bf16 contraction (pointwise.cu).  Its bf16 output, read back from the workspace,
    must be the exact convolution rounded to bf16: at most one bf16 ulp apart, and exactly equal for most
    elements.  Shapes off the 64-frame / 32-bin tile included; log-mel-like magnitudes (tens of dB)."""
    import torch.nn.functional as F
    mt = "cnn_rnn"
    sd = synth.synth_state_dict(mt, n_mels, 128, 1, seed=4)
    m = TranscriptionModel(model_type=mt, n_mels=n_mels, hidden_size=128, num_layers=1, device=DEV).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(n_mels + T)
    x = (torch.randn(B, 1, n_mels, T, generator=g) * 20.0 - 10.0).to(DEV)
    m(x)
    F1 = n_mels // 2
    got = m.workspace_tensor("act1", B, T, torch.bfloat16, F1 * 32).float().view(B, T, F1, 32)
    w = m.packed_tensor("conv1.w", torch.float32).view(32, 1, 3, 3)
    bias = m.packed_tensor("conv1.b", torch.float32)
    ref = F.max_pool2d(F.conv2d(x.double(), w.double(), bias.double(), padding=1).relu(), (2, 1))     # (B, 32, F1, T)
    ref = ref.permute(0, 3, 2, 1).float()
    # size of the terms that were summed (the split-bf16 contraction is exact to 2^-16 of THAT, not of the result)
    mag = F.max_pool2d(F.conv2d(x.double().abs(), w.double().abs(), bias.double().abs(), padding=1), (2, 1)).permute(0, 3, 2, 1).float()
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + mag * 2.0 ** -14      # one bf16 ulp of the value (2^-8 .. 2^-7) + the contraction's own error
    assert bool((err <= tol).all()), float((err / tol).max())
    exact = (got == ref.to(torch.bfloat16).float()).float().mean().item()
    assert exact > 0.85, exact                          # (a 2^-16 error flips a bf16 rounding in ~1 % x cancellation of the cases)


def test_three_heads_decode_to_notes_end_to_end():
    """audio -> log-mel -> CNNRNNModelLarge (all heads) -> onset/offset-aware notes: the decoder sees the same thresholded
    rolls the oracle sees (sigmoid and strict float32 compare on the GPU == numpy on the downloaded logits up to cells that
    sit exactly on the threshold, which random-init logits do not), so the note lists must be equal."""
    from oracle import notes as onotes
    fe = pipeline.Frontend.get(device=DEV)
    wav = torch.from_numpy(synth.piano_chord_batch([1, 2], n_samples=48000)).to(DEV)
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=128, num_layers=1, device=DEV).eval()
    m.load_state_dict(synth.synth_state_dict("cnn_rnn_large", 320, 128, 1, seed=6))
    heads = m(fe.logmel(wav, defer_floor=True), return_all_heads=True)
    got = pipeline.extract_notes_onset_aware(heads["frame"], heads["onset"], heads["offset"], 0.5, 0.5, 0.5)
    roll = lambda k: np.concatenate(list((torch.sigmoid(heads[k]).cpu().numpy() > np.float32(0.5)).astype(np.float32)), axis=1)
    want = onotes.group_notes_onset_aware(roll("frame"), roll("onset"), roll("offset"))
    assert len(want) > 0 and np.array_equal(got, want)
    nl = pipeline.NoteList(got, fs=16000 / 512)
    assert len(nl) == len(got)


def _load_case(path):
    g = np.load(path)
    n_mels, H, L, B, T, attn, heads, seed, xseed = [int(v) for v in g["cfg"]]
    mt = str(g["model_type"])
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=seed, use_attention=bool(attn), use_onset_offset_heads=bool(heads))
    return g, mt, n_mels, H, L, bool(attn), bool(heads), sd


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*_*.npz"))),
                         ids=lambda p: os.path.basename(p)[6:-4])
def test_model_matches_reference_golden(path):
    g, mt, n_mels, H, L, attn, heads, sd = _load_case(path)
    m = TranscriptionModel(model_type=mt, n_mels=n_mels, hidden_size=H, num_layers=L, dropout=0.2, device=DEV,
                           use_attention=attn, use_onset_offset_heads=heads)
    m.load_state_dict(sd, strict=True)          # reference .pth key set
    m.eval()
    x = torch.from_numpy(g["x"]).to(DEV)
    out = m(x)
    assert out.shape == g["frame"].shape and out.dtype == torch.float32 and out.device.type == "cuda"
    got = {"frame": out}
    if mt.endswith("large") and heads:
        allh = m(x, return_all_heads=True)
        assert sorted(allh) == ["frame", "offset", "onset"]
        assert torch.equal(allh["frame"], out)
        got = allh
    for k, v in got.items():
        ref = torch.from_numpy(g[k])
        dl = (v.cpu() - ref).abs()
        dp = (torch.sigmoid(v.cpu()) - torch.sigmoid(ref)).abs()
        _report(f"fast.{os.path.basename(path)[6:-4]}.{k}", prob_max=dp.max(), prob_mean=dp.mean(), logit_max=dl.max())
        assert dl.max() < LOGIT_MAX and dp.max() < PROB_MAX and dp.mean() < PROB_MEAN, (k, dl.max(), dp.max(), dp.mean())
    # kernels vs a CPU emulation of the same bf16 roundings: isolates kernel error from the bf16 floor
    from music_transcription_b200.packing import pack_state_dict
    from tests.emulate import emu_forward
    emu = emu_forward(pack_state_dict(sd, mt, n_mels, H, L, attn, heads), torch.from_numpy(g["x"]), mt, n_mels, H, L,
                      attn, heads, bf16_acts=True)
    for k, v in got.items():
        dl = (v.cpu() - emu[k]).abs()
        dp = (torch.sigmoid(v.cpu()) - torch.sigmoid(emu[k])).abs()
        assert dl.max() < 5e-2 and dp.max() < 1e-2 and dp.mean() < 1e-3, ("emu", k, dl.max(), dp.max(), dp.mean())
    pred = m.predict(x, threshold=0.5)
    assert pred.shape == g["pred"].shape and set(np.unique(pred.cpu().numpy())) <= {0.0, 1.0}
    flips = (pred.cpu().numpy() != g["pred"]).mean()
    _report(f"fast.{os.path.basename(path)[6:-4]}.pred", flips=flips)
    assert flips < 0.009                          # stress checkpoints: measured 0.19-0.68 % (profiles/r2_parity.md)


@pytest.mark.parametrize("mt,n_mels,H,L,attn,heads", [
    ("cnn_rnn_large", 64, 128, 2, True, True), ("cnn_rnn_large", 37, 128, 1, False, True), ("cnn_rnn_large", 64, 256, 1, True, False),
    ("cnn_rnn", 64, 128, 2, True, True), ("cnn_rnn", 37, 256, 1, True, True)])
@pytest.mark.parametrize("precision", ["fast", "precise"])
def test_library_packing_is_bitwise_the_python_packing(mt, n_mels, H, L, attn, heads, precision):
    """amt_model_load (CUDA kernels inside the library: BN fold in double, layouts, permutations, bf16 / split-bf16)
    must produce exactly the tensors packing.pack_state_dict states in torch -- every packed tensor, bit for bit."""
    from music_transcription_b200.packing import pack_state_dict
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=9, use_attention=attn, use_onset_offset_heads=heads)
    m = TranscriptionModel(mt, n_mels=n_mels, hidden_size=H, num_layers=L, device=DEV, use_attention=attn,
                           use_onset_offset_heads=heads, precision=precision)
    m.load_state_dict(sd, strict=True)
    m(torch.zeros(1, 1, n_mels, 8, device=DEV))                       # packs
    want = pack_state_dict(sd, mt, n_mels, H, L, attn, heads, precise=precision == "precise")
    for name, t in want.items():
        got = m.packed_tensor(name, t.dtype).cpu().view(t.shape)
        assert torch.equal(got.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32),
                           t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32)), (name, (got.float() - t.float()).abs().max())


def test_model_load_reports_missing_and_missized_tensors():
    import ctypes as C
    from music_transcription_b200 import _lib
    L = _lib.lib()
    cfg = _lib.ModelConfig(0, 64, 128, 1, 8, 1, 1, 0)
    h = C.c_void_p()
    _lib.check(L.amt_model_create(C.byref(cfg), C.byref(h)))
    sd = {k: v.to(DEV).float().contiguous() for k, v in synth.synth_state_dict("cnn_rnn", 64, 128, 1, seed=1).items() if v.is_floating_point()}

    def load(d):
        n = len(d)
        return L.amt_model_load(h, (C.c_char_p * n)(*[k.encode() for k in d]), (C.c_void_p * n)(*[t.data_ptr() for t in d.values()]),
                                (C.c_int64 * n)(*[t.numel() for t in d.values()]), n, _lib.stream_ptr(torch.device(DEV)))
    missing = {k: v for k, v in sd.items() if k != "model.fc.bias"}
    assert load(missing) == _lib.AMT_ERR_STATE and b"model.fc.bias" in L.amt_last_error()
    bad = dict(sd)
    bad["model.cnn.4.weight"] = bad["model.cnn.4.weight"][:32].contiguous()
    assert load(bad) == _lib.AMT_ERR_STATE and b"model.cnn.4.weight" in L.amt_last_error()
    assert load(sd) == 0                                               # and the handle is usable after a failed attempt
    L.amt_model_destroy(h)


def test_stage_profile_and_in_flight_query():
    m = TranscriptionModel(model_type="cnn_rnn", n_mels=64, hidden_size=128, num_layers=1, device=DEV)
    m.load_state_dict(synth.synth_state_dict("cnn_rnn", 64, 128, 1, seed=3))
    m.eval()
    x = torch.randn(2, 1, 64, 50, device=DEV)
    ref = m(x).clone()
    m.profile(True)
    out = m(x)
    torch.cuda.synchronize()
    assert m.profile_in_flight() is None                 # everything launched has completed
    stages = m.profile_read()
    assert torch.equal(out, ref)
    names = [s[0] for s in stages]
    assert "conv1" in names and any(n.endswith(".rec") for n in names)
    assert all(ms > 0 and n == 1 for _, ms, n in stages)
    m.profile(False)


def test_compute_loss_matches_reference_golden_and_oracle():
    """amt_bce_loss (one CUDA pass, fp64 accumulation) vs the reference's own compute_loss values
    (tests/golden/loss_reference.npz) and the oracle restatement: relative 2e-6."""
    from oracle import losses as olosses
    from tests.test_oracle import _loss_cases
    m = TranscriptionModel(model_type="cnn_rnn", n_mels=64, hidden_size=128, num_layers=1, device=DEV)
    for name, logits, roll, lengths, ref in _loss_cases():
        dl = {k: v.to(DEV) for k, v in logits.items()} if isinstance(logits, dict) else logits.to(DEV)
        got = m.compute_loss(dl, roll.to(DEV), None if lengths is None else lengths.to(DEV))
        assert got.shape == () and got.dtype == torch.float32 and got.is_cuda
        assert abs(float(got) - ref) <= 2e-6 * max(1.0, abs(ref)), (name, float(got), ref)
    # canonical size, three heads, ragged lengths: against the oracle
    g = torch.Generator().manual_seed(5)
    heads = {k: torch.randn(16, 88, 938, generator=g) * 2 for k in ("frame", "onset", "offset")}
    roll = (torch.rand(16, 88, 938, generator=g) < 0.05).float()
    lengths = torch.tensor([938, 937, 469, 1] * 4)
    ref = float(olosses.compute_loss(heads, roll, lengths))
    got = float(m.compute_loss({k: v.to(DEV) for k, v in heads.items()}, roll.to(DEV), lengths))
    assert abs(got - ref) <= 2e-6 * ref
    parts = m.last_loss_parts.cpu()
    assert abs(float(0.5 * parts[0] + 0.25 * parts[1] + 0.25 * parts[2]) - got) < 1e-6
    with pytest.raises(ValueError):
        m.compute_loss(heads["frame"].to(DEV), roll[:, :40].to(DEV))
    with pytest.raises(Exception):
        m.compute_loss(heads["frame"], roll)          # CPU tensors: no fallback


def test_model_rejects_bad_inputs():
    with pytest.raises(ValueError):
        TranscriptionModel(model_type="nope", device=DEV)
    m = TranscriptionModel(model_type="cnn_rnn", n_mels=64, hidden_size=128, num_layers=1, device=DEV)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 64, 0, device=DEV))
    with pytest.raises(Exception):
        m(torch.zeros(1, 1, 64, 10))          # CPU tensor: no fallback


def test_canonical_large_model_vs_oracle_full_chunk():
    """CNNRNNModelLarge at the reference's canonical config (n_mels 320, hidden 512, 3 layers), 30-s
    chunks of real log-mel; oracle = fp32 PyTorch restatement on the CPU.
    (a) default-init-scale weights (gain 1/sqrt(3) == torch's kaiming/LSTM default variance): the stated
        product tolerance of the fast mode, probabilities max-abs <= 2e-3 (the bf16 operand floor, tests/attribution.py);
    (b) stress weights (unit gain, logits +-3, chaotic LSTM dynamics): the kernels must match a CPU
        emulation of the same bf16 roundings to 5e-2 / 1e-2 / 1e-3, and stay within 1.25x of that
        emulation's own distance to the fp32 oracle (chunk 3 has a measured bf16 floor of 0.35 logits)."""
    from music_transcription_b200.packing import pack_state_dict
    from oracle import model as omodel
    from tests.emulate import emu_forward
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    wav = torch.from_numpy(synth.piano_chord_batch([0, 3])).to(DEV)
    mel = pipeline.Frontend.get(device=DEV).logmel(wav)
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, dropout=0.2, device=DEV)
    # (a)
    sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5)
    m.load_state_dict(sd)
    out = m(mel, return_all_heads=True)
    ref = omodel.large_forward(sd, mel.cpu(), 512, 3, return_all_heads=True)
    for k in ("frame", "onset", "offset"):
        dl = (out[k].cpu() - ref[k]).abs()
        dp = (torch.sigmoid(out[k].cpu()) - torch.sigmoid(ref[k])).abs()
        _report(f"canon_vs_oracle.{k}", prob_max=dp.max(), prob_mean=dp.mean(), logit_max=dl.max())
        assert dp.max() < CANON_PROB_MAX and dp.mean() < CANON_PROB_MEAN, ("default-init", k, dl.max(), dp.max(), dp.mean())
    # (b)
    sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1)
    m.load_state_dict(sd)           # in-place update -> repacked on the next call
    out = m(mel, return_all_heads=True)
    ref = omodel.large_forward(sd, mel.cpu(), 512, 3, return_all_heads=True)
    emu = emu_forward(pack_state_dict(sd, "cnn_rnn_large", 320, 512, 3), mel.cpu(), "cnn_rnn_large", 320, 512, 3)
    for k in ("frame", "onset", "offset"):
        g = out[k].cpu()
        de = (g - emu[k]).abs()
        dpe = (torch.sigmoid(g) - torch.sigmoid(emu[k])).abs()
        assert de.max() < 5e-2 and dpe.max() < 1e-2 and dpe.mean() < 1e-3, ("emu", k, de.max(), dpe.max(), dpe.mean())
        floor = (emu[k] - ref[k]).abs().max()
        assert (g - ref[k]).abs().max() < 1.25 * floor + 1e-2, ("stress", k, (g - ref[k]).abs().max(), floor)
    # batch independence: chunk 1 alone gives the same logits as inside the batch
    solo = m(mel[1:2])
    assert (solo[0] - out["frame"][1]).abs().max() < 1e-5


def _report(name, **vals):
    """Measured parity numbers -> gpurun_out/parity_report.jsonl (scratch; summarised under profiles/ by hand)."""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: float(v) for k, v in vals.items()}}) + "\n")


CANON_PROB_MAX, CANON_PROB_MEAN, CANON_FLIPS = 2e-3, 3.5e-4, 4e-3    # fast (bf16) mode; measured 1.55e-3 / 2.6e-4 / 2.8e-3 (profiles/r2_parity.md)


@pytest.mark.parametrize("name", ["large", "small"])
def test_canonical_shape_goldens_from_the_real_reference(name):
    """tests/golden/canon_*.npz: outputs of the UNMODIFIED reference modules at the canonical shapes (n_mels 320,
    hidden 512, 3 layers, T 938: the 16-CTA-cluster LSTM, head-dim-192 attention, K = 10240 projection) on chord
    log-mel.  CNNRNNModelLarge (three heads, 2 chunks) and CNNRNNModel 36 M (BASELINE configs[0]'s model)."""
    g = np.load(os.path.join(GOLDEN, f"canon_{name}.npz"))
    n_mels, H, L, B, T, seed = [int(v) for v in g["cfg"]]
    mt = str(g["model_type"])
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=seed, gain=float(g["gain"]))
    m = TranscriptionModel(mt, n_mels=n_mels, hidden_size=H, num_layers=L, dropout=0.2, device=DEV)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = torch.from_numpy(g["x"].astype(np.float32)).to(DEV)
    out = m(x, return_all_heads=True)
    out = out if isinstance(out, dict) else {"frame": out}
    assert sorted(out) == (["frame", "offset", "onset"] if name == "large" else ["frame"])
    for k, v in out.items():
        ref = torch.from_numpy(g[k])
        assert v.shape == ref.shape == (B, 88, T)
        dp = (torch.sigmoid(v.cpu()) - torch.sigmoid(ref)).abs()
        flips = ((v.cpu() > 0) != (ref > 0)).float().mean()
        _report(f"canon_{name}.{k}", prob_max=dp.max(), prob_mean=dp.mean(), logit_max=(v.cpu() - ref).abs().max(), flips=flips)
        assert dp.max() < CANON_PROB_MAX and dp.mean() < CANON_PROB_MEAN and flips < CANON_FLIPS, (k, dp.max(), dp.mean(), flips)


PRECISE_PROB_MAX, PRECISE_PROB_MEAN, PRECISE_FLIPS = 4e-4, 8e-5, 1e-3     # precise (split-bf16) mode; measured: profiles/r2_parity.md


@pytest.mark.parametrize("name", ["large", "small"])
def test_precise_mode_matches_the_real_reference_at_canonical_shapes(name):
    """precision="precise": split-bf16 operands (three MMA products per contraction).  Against the outputs of the
    unmodified fp32 reference at the canonical shapes: probabilities within 4e-4 (north_star's example tolerance is
    1e-3), fewer than 0.1 % thresholded cells differ.  What remains is the recurrence (bf16 W_hh / h feedback,
    tanh.approx) and the attention core (bf16 q, k, v, P) -- tests/attribution.py predicts 2e-4."""
    g = np.load(os.path.join(GOLDEN, f"canon_{name}.npz"))
    n_mels, H, L, B, T, seed = [int(v) for v in g["cfg"]]
    mt = str(g["model_type"])
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=seed, gain=float(g["gain"]))
    m = TranscriptionModel(mt, n_mels=n_mels, hidden_size=H, num_layers=L, dropout=0.2, device=DEV, precision="precise")
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = torch.from_numpy(g["x"].astype(np.float32)).to(DEV)
    out = m(x, return_all_heads=True)
    out = out if isinstance(out, dict) else {"frame": out}
    for k, v in out.items():
        ref = torch.from_numpy(g[k])
        dp = (torch.sigmoid(v.cpu()) - torch.sigmoid(ref)).abs()
        flips = ((v.cpu() > 0) != (ref > 0)).float().mean()
        _report(f"precise.canon_{name}.{k}", prob_max=dp.max(), prob_mean=dp.mean(), logit_max=(v.cpu() - ref).abs().max(), flips=flips)
        assert dp.max() < PRECISE_PROB_MAX and dp.mean() < PRECISE_PROB_MEAN and flips < PRECISE_FLIPS, (k, dp.max(), dp.mean(), flips)
    # switching the same object back to fast mode re-packs and reproduces the fast result
    fast = TranscriptionModel(mt, n_mels=n_mels, hidden_size=H, num_layers=L, dropout=0.2, device=DEV)
    fast.load_state_dict(sd)
    assert torch.equal(m.set_precision("fast")(x), fast(x))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*_*.npz"))),
                         ids=lambda p: os.path.basename(p)[6:-4])
def test_precise_mode_on_the_small_reference_goldens(path):
    """The six reference-run fixtures (unit-gain stress weights, odd n_mels, no attention / no heads) in precise mode:
    one order of magnitude inside the fast-mode tolerance."""
    g, mt, n_mels, H, L, attn, heads, sd = _load_case(path)
    m = TranscriptionModel(model_type=mt, n_mels=n_mels, hidden_size=H, num_layers=L, dropout=0.2, device=DEV,
                           use_attention=attn, use_onset_offset_heads=heads, precision="precise")
    m.load_state_dict(sd, strict=True)
    x = torch.from_numpy(g["x"]).to(DEV)
    got = m(x, return_all_heads=True) if (mt.endswith("large") and heads) else {"frame": m(x)}
    for k, v in got.items():
        ref = torch.from_numpy(g[k])
        dl = (v.cpu() - ref).abs()
        dp = (torch.sigmoid(v.cpu()) - torch.sigmoid(ref)).abs()
        _report(f"precise.{os.path.basename(path)[6:-4]}.{k}", prob_max=dp.max(), prob_mean=dp.mean(), logit_max=dl.max())
        assert dl.max() < LOGIT_MAX / 8 and dp.max() < PROB_MAX / 8 and dp.mean() < PROB_MEAN / 8, (k, dl.max(), dp.max(), dp.mean())


@pytest.mark.parametrize("B", [16, 64])
def test_large_model_batch_16_and_64_invariance_and_oracle(B):
    """BASELINE configs[2] (batch 16) and the bench batch (64) on the north-star model: every chunk's logits are
    bitwise what the same chunk gives alone (B = 1), and 4 of the chunks agree with the fp32 oracle."""
    from oracle import model as omodel
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5)
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, dropout=0.2, device=DEV)
    m.load_state_dict(sd)
    wav = torch.from_numpy(synth.piano_chord_batch(range(16))).to(DEV)
    if B > 16:
        wav = torch.cat([wav * (1.0 - 0.05 * j) for j in range(B // 16)])
    mel = pipeline.Frontend.get(device=DEV).logmel(wav)
    out = m(mel, return_all_heads=True)
    picks = [0, 5, B // 2 + 1, B - 1]
    for i in picks:
        solo = m(mel[i:i + 1], return_all_heads=True)
        for k in ("frame", "onset", "offset"):
            assert torch.equal(solo[k][0], out[k][i]), (B, i, k, (solo[k][0] - out[k][i]).abs().max().item())
    ref = omodel.large_forward(sd, mel[picks].cpu(), 512, 3, return_all_heads=True)
    for k in ("frame", "onset", "offset"):
        dp = (torch.sigmoid(out[k][picks].cpu()) - torch.sigmoid(ref[k])).abs()
        _report(f"large_B{B}.{k}", prob_max=dp.max(), prob_mean=dp.mean())
        assert dp.max() < CANON_PROB_MAX and dp.mean() < CANON_PROB_MEAN, (B, k, dp.max(), dp.mean())


def test_two_hour_recording_240_chunks_on_the_large_model():
    """BASELINE configs[3] on the north-star model: 240 x 30-s chunks through CNNRNNModelLarge (320/512/3).  The
    oracle cannot run at this size, so: (i) the note list does not depend on the batching (240 at once vs 64-chunk
    batches), (ii) it equals the oracle's grouping of the SAME GPU probability roll, seams included, (iii) a chunk's
    logits are bitwise independent of its batch and position."""
    from oracle import notes as onotes
    sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5)
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, dropout=0.2, device=DEV)
    m.load_state_dict(sd)
    base = synth.cheap_wave_batch(8, 480000, seed=7)
    wav = torch.empty(240, 480000, device=DEV)
    for i in range(240):
        wav[i] = base[i % 8].to(DEV) * (1.0 - 0.003 * (i // 8))
    notes_a, probs = pipeline.transcribe_chunks(m, wav, threshold=0.5, batch=240, return_probs=True)
    notes_b, _ = pipeline.transcribe_chunks(m, wav, threshold=0.5, batch=64)
    assert len(notes_a) > 0 and np.array_equal(notes_a, notes_b)
    p = probs.cpu().numpy()
    want = onotes.group_notes(onotes.combine_piano_rolls([onotes.threshold_roll(p[i], 0.5) for i in range(240)]))
    assert np.array_equal(notes_a, want)
    fe = pipeline.Frontend.get(device=DEV)
    solo = m(fe.logmel(wav[100:101]))
    batch = m(fe.logmel(wav[96:112]))
    assert torch.equal(solo[0], batch[4])


def test_transcribe_chunks_end_to_end_notes_match_oracle_on_same_probs():
    from oracle import notes as onotes
    sd = synth.synth_state_dict("cnn_rnn", 320, 128, 1, seed=2, gain=2.0)
    m = TranscriptionModel("cnn_rnn", n_mels=320, hidden_size=128, num_layers=1, device=DEV)
    m.load_state_dict(sd)
    wav = torch.from_numpy(synth.piano_chord_batch([0, 1, 2], n_samples=160000)).to(DEV)
    notes, probs = pipeline.transcribe_chunks(m, wav, threshold=0.5, batch=2, return_probs=True)
    p = probs.cpu().numpy()
    rolls = [onotes.threshold_roll(p[i], 0.5) for i in range(3)]
    want = onotes.group_notes(onotes.combine_piano_rolls(rolls))
    assert np.array_equal(notes, want)
    nl = pipeline.pianoroll_to_midi(onotes.combine_piano_rolls(rolls), fs=16000 / 512)
    assert len(nl) == len(want) and all(n.velocity == 100 for n in nl.instruments[0].notes)


def test_two_hour_recording_240_chunks_batch_invariance_and_notes():
    """BASELINE configs[3] shape: 240 x 30-s chunks (full T = 938).  The oracle cannot run at this size, so the
    test uses size-independent properties: (i) the note list does not depend on how the chunks are batched
    (240 at once vs 64 + 64 + 64 + 48), (ii) it equals the oracle's grouping of the SAME GPU probability roll,
    seams included, (iii) per-chunk logits are bitwise independent of the batch position."""
    from oracle import notes as onotes
    sd = synth.synth_state_dict("cnn_rnn", 320, 128, 1, seed=4, gain=2.0)
    m = TranscriptionModel("cnn_rnn", n_mels=320, hidden_size=128, num_layers=1, device=DEV)
    m.load_state_dict(sd)
    base = synth.cheap_wave_batch(8, 480000, seed=7)
    wav = torch.empty(240, 480000, device=DEV)
    for i in range(240):
        wav[i] = base[i % 8].to(DEV) * (1.0 - 0.003 * (i // 8))
    notes_a, probs = pipeline.transcribe_chunks(m, wav, threshold=0.5, batch=240, return_probs=True)
    notes_b, _ = pipeline.transcribe_chunks(m, wav, threshold=0.5, batch=64)
    assert np.array_equal(notes_a, notes_b)
    p = probs.cpu().numpy()
    want = onotes.group_notes(onotes.combine_piano_rolls([onotes.threshold_roll(p[i], 0.5) for i in range(240)]))
    assert np.array_equal(notes_a, want)
    fe = pipeline.Frontend.get(device=DEV)
    solo = m(fe.logmel(wav[100:101]))
    batch = m(fe.logmel(wav[96:112]))
    assert torch.equal(solo[0], batch[4])


def test_two_forwards_in_flight_on_two_streams_are_bitwise_one_stream():
    """One model object used from two CUDA streams at once (what StreamingTranscriber(lanes=2) does): every stream gets its
    own workspace, the packed weights are shared, and the results are bitwise those of a single stream -- on the
    north-star model, whose recurrences then overlap the other stream's tensor kernels."""
    sd = synth.synth_state_dict("cnn_rnn_large", 320, 512, 3, seed=1, gain=3 ** -0.5)
    m = TranscriptionModel("cnn_rnn_large", n_mels=320, hidden_size=512, num_layers=3, dropout=0.2, device=DEV).eval()
    m.load_state_dict(sd)
    fe = pipeline.Frontend.get(device=DEV)
    wavs = [torch.from_numpy(synth.piano_chord_batch(range(4 * k, 4 * k + 4))).to(DEV) for k in range(2)]
    ref = [m(fe.logmel(w)).clone() for w in wavs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(DEV), torch.cuda.Stream(DEV)]
    outs = []
    for rep in range(6):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                outs.append((k, m(fe.logmel(wavs[k]))))
    torch.cuda.synchronize()
    assert len(m._workspaces) >= 3                              # the default stream's and one per lane
    for k, o in outs:
        assert torch.equal(o, ref[k])


@pytest.mark.parametrize("input_format,roll_format,lanes", [("f32", "f32", 1), ("pcm16", "bits", 1), ("f32", "bits", 2), ("pcm16", "bits", 2)])
def test_streaming_transcriber_equals_batch_by_batch_path(input_format, roll_format, lanes):
    """Overlapped H2D / compute / D2H (3 streams, 2 slots) must return exactly what the plain per-batch path does,
    including a short last batch and slot reuse (5 batches over 2 slots) -- for float32 and 16-bit PCM input (the
    on-device conversion is bit-identical to sample / 32768 on the host) and for the float and the bit-packed roll."""
    sd = synth.synth_state_dict("cnn_rnn", 320, 128, 1, seed=5, gain=2.0)
    m = TranscriptionModel("cnn_rnn", n_mels=320, hidden_size=128, num_layers=1, device=DEV)
    m.load_state_dict(sd)
    base = synth.cheap_wave_batch(8, 480000, seed=11)
    batches = [(base[:4] * (1.0 - 0.1 * k)).contiguous() for k in range(4)] + [base[4:6].contiguous()]
    if input_format == "pcm16":
        pcm = [torch.clamp(torch.round(b * 32768.0), -32768, 32767).to(torch.int16).pin_memory() for b in batches]
        batches = [p.float() / 32768.0 for p in pcm]          # what a host-side decode of the same file gives
        feed = pcm
    else:
        feed = [b.pin_memory() for b in batches]
    st = pipeline.StreamingTranscriber(m, 4, 480000, 0.5, input_format=input_format, roll_format=roll_format, lanes=lanes)
    got = [(r.clone(), n.copy()) for r, n in st.run(feed)]
    assert len(got) == 5
    for hb, (roll, notes) in zip(batches, got):
        want_notes, probs = pipeline.transcribe_chunks(m, hb.to(DEV), threshold=0.5, batch=4, return_probs=True)
        assert np.array_equal(notes, want_notes)
        want_roll = (probs > 0.5).float().cpu()
        if roll_format == "bits":
            assert roll.dtype == torch.int32 and roll.shape == (hb.shape[0], 88, (st.T + 31) // 32)
            assert np.array_equal(pipeline.unpack_roll(roll, st.T), want_roll.numpy())
        else:
            assert torch.equal(roll, want_roll)
    # copy=True hands out private copies: collecting the generator must not alias the two pinned slots
    st2 = pipeline.StreamingTranscriber(m, 4, 480000, 0.5, input_format=input_format, roll_format=roll_format, copy=True, lanes=lanes)
    kept = list(st2.run(feed))
    for (r0, n0), (r1, n1) in zip(got, kept):
        assert torch.equal(r0, r1) and np.array_equal(n0, n1)
