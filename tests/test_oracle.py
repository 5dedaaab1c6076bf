"""CPU: the oracle against the golden vectors produced by the real reference
(oracle/make_golden.py) and against sklearn / closed-form properties."""
import glob
import os

import numpy as np
import pytest
import torch

from music_transcription_b200 import synth
from oracle import f1 as of1
from oracle import frontend as ofe
from oracle import model as omodel
from oracle import notes as onotes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*_*.npz"))),
                         ids=lambda p: os.path.basename(p)[6:-4])
def test_model_oracle_matches_reference_outputs(path):
    g = np.load(path)
    n_mels, H, L, B, T, attn, heads, seed, xseed = [int(v) for v in g["cfg"]]
    mt = str(g["model_type"])
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=seed, use_attention=bool(attn),
                                use_onset_offset_heads=bool(heads))
    assert list(sd.keys()) == [str(k) for k in g["keys"]]          # checkpoint key parity
    x = torch.from_numpy(g["x"])
    assert torch.equal(x, synth.synth_logmel(B, n_mels, T, seed=xseed))
    torch.set_num_threads(4)
    out = omodel.forward(sd, x, mt, H, L, bool(attn), bool(heads), return_all_heads=True)
    if isinstance(out, dict):
        for k in ("frame", "onset", "offset"):
            np.testing.assert_allclose(out[k].numpy(), g[k], atol=2e-5, rtol=0)
        frame = out["frame"]
    else:
        frame = out
        np.testing.assert_allclose(frame.numpy(), g["frame"], atol=2e-5, rtol=0)
    assert frame.shape == (B, 88, T)
    pred = omodel.predict(sd, x, mt, H, L, 0.5, use_attention=bool(attn), use_onset_offset_heads=bool(heads))
    assert (pred.numpy() != g["pred"]).mean() < 1e-3


@pytest.mark.parametrize("name", ["large", "small"])
def test_model_oracle_matches_reference_at_canonical_shapes(name):
    """tests/golden/canon_*.npz = the real reference modules at n_mels 320 / hidden 512 / 3 layers / T 938 on chord
    log-mel (oracle/make_golden.py canonical): the port must reproduce them to fp32 rounding."""
    g = np.load(os.path.join(GOLDEN, f"canon_{name}.npz"))
    n_mels, H, L, B, T, seed = [int(v) for v in g["cfg"]]
    mt = str(g["model_type"])
    sd = synth.synth_state_dict(mt, n_mels, H, L, seed=seed, gain=float(g["gain"]))
    x = torch.from_numpy(g["x"].astype(np.float32))
    assert x.shape == (B, 1, n_mels, T)
    # the stored input IS the oracle frontend's log-mel of the chord chunks, on the float16 grid
    want = ofe.logmel(synth.piano_chord(int(g["chunks"][0]))).astype(np.float16).astype(np.float32)
    assert np.abs(want - g["x"][0, 0].astype(np.float32)).max() <= 0.0625      # one fp16 ulp at |x| < 64 (libm differences)
    torch.set_num_threads(os.cpu_count() or 4)
    out = omodel.forward(sd, x, mt, H, L, return_all_heads=True)
    out = out if isinstance(out, dict) else {"frame": out}
    for k, v in out.items():
        np.testing.assert_allclose(v.numpy(), g[k], atol=5e-5, rtol=0)


def test_reference_rejects_zero_length_input():
    g = np.load(os.path.join(GOLDEN, "model_T0.npz"))
    assert "Kernel size" in str(g["raised"])      # the T==0 guard of the reference is unreachable


def test_notes_oracle_matches_reference_pianoroll_to_midi():
    g = np.load(os.path.join(GOLDEN, "notes_reference.npz"))
    fs = 16000 / 512
    for name in ("random", "sparse", "full", "empty", "edges", "seam"):
        roll = g[name + "_roll"].astype(np.float32)
        notes = onotes.group_notes(roll)
        ev = onotes.notes_to_events(notes, fs)
        assert len(ev) == len(g[name + "_pitch"]), name
        if len(ev):
            assert np.array_equal([e[0] for e in ev], g[name + "_pitch"])
            assert np.array_equal(np.array([e[1] for e in ev]), g[name + "_start"])   # bit-exact float64
            assert np.array_equal(np.array([e[2] for e in ev]), g[name + "_end"])
            assert set(g[name + "_vel"]) == {100}
    comb = onotes.combine_piano_rolls([g["seam_a"], g["seam_b"]])
    assert np.array_equal(comb, g["seam_roll"])
    seam = onotes.group_notes(comb.astype(np.float32))
    assert any(p == 5 and s == 930 and e == 945 for p, s, e in seam)        # merged across the chunk seam


def test_onset_aware_decoding_rule_on_hand_cases():
    """The onset / offset-aware decoder has no counterpart in the reference (its inference drops those heads): the rule is
    ours (amt.h), this pins its CPU definition on cases whose answer is obvious."""
    from oracle.notes import group_notes_onset_aware as dec
    F = np.array([[0, 1, 1, 1, 0, 1, 1, 1, 1, 0]])
    ON = np.array([[0, 1, 0, 0, 0, 0, 1, 1, 0, 0]])
    OFF = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 1, 0]])
    assert dec(F, ON).tolist() == [[0, 1, 4], [0, 6, 9]]            # frame 5 sounds but no onset opened it
    assert dec(F, ON, OFF).tolist() == [[0, 1, 4], [0, 6, 8]]       # the offset head ends the second note early
    # re-strike inside a sounding note; an onset on a silent frame still sounds; open at the end
    F = np.array([[1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 0]])
    ON = np.array([[1, 0, 0, 1, 0, 0], [0, 0, 1, 1, 0, 1]])
    assert dec(F, ON).tolist() == [[0, 0, 3], [0, 3, 6], [1, 2, 4], [1, 5, 6]]
    assert dec(np.zeros((2, 5)), np.zeros((2, 5))).shape == (0, 3)


def test_onset_aware_decoding_loop_equals_a_vectorised_restatement():
    """The loop in oracle/notes.py is the definition; the same rule written with array operations (starts = rising onset
    edges, boundaries = silence | start | offset, each start paired with the next boundary by searchsorted) must give the
    same notes on random rolls -- two independent writings of the rule the GPU kernel is held to."""
    from oracle.notes import group_notes_onset_aware as dec
    rng = np.random.default_rng(11)
    for trial in range(20):
        P, N = int(rng.integers(1, 6)), int(rng.integers(1, 300))
        F = rng.random((P, N)) < rng.uniform(0.1, 0.9)
        ON = rng.random((P, N)) < rng.uniform(0.02, 0.6)
        OFF = (rng.random((P, N)) < rng.uniform(0.0, 0.3)) if trial % 2 else None
        rows = []
        for p in range(P):
            prev = np.concatenate([[False], ON[p, :-1]])
            S = ON[p] & ~prev
            B = ~(F[p] | ON[p]) | S | (OFF[p] if OFF is not None else False)
            starts, bounds = np.flatnonzero(S), np.flatnonzero(B)
            k = np.searchsorted(bounds, starts, side="right")             # first boundary strictly after the start
            ends = np.where(k < len(bounds), bounds[np.minimum(k, len(bounds) - 1)] if len(bounds) else N, N)
            rows += [(p, int(a), int(b)) for a, b in zip(starts, ends)]
        want = np.asarray(rows, dtype=np.int32).reshape(-1, 3)
        assert np.array_equal(dec(F, ON, OFF), want), trial


def test_threshold_is_strict_float32_compare():
    p = np.array([[np.float32(0.1), np.nextafter(np.float32(0.1), np.float32(1))]], dtype=np.float32)
    assert onotes.threshold_roll(p, 0.1).tolist() == [[0.0, 1.0]]
    assert (torch.from_numpy(p) > 0.1).float().numpy().tolist() == [[0.0, 1.0]]


def test_f1_oracle_matches_reference_evaluate():
    g = np.load(os.path.join(GOLDEN, "f1_reference.npz"))
    probs, rolls, lengths = g["probs"], g["rolls"].astype(np.float32), g["lengths"]

    def mean_at(t):
        return of1.mean_f1(of1.counts_grid(probs, rolls, lengths, [t])[:, 0])

    for t, want in zip(g["at_t"], g["at_f1"]):
        assert mean_at(float(t)) == pytest.approx(float(want), abs=1e-15)
    best_t, best_f1, visited = of1.threshold_walk(mean_at)
    assert best_t == float(g["best_t"]) and best_f1 == pytest.approx(float(g["best_f1"]), abs=1e-15)
    # default schedule with an interior optimum: 10+9+9+9 thresholds in 4 rounds
    _, _, v2 = of1.threshold_walk(lambda t: -abs(t - 0.52))
    assert len(v2) == 37


def test_f1_counts_match_sklearn():
    from sklearn.metrics import confusion_matrix, f1_score
    for i in range(4):
        p = synth.planted_probs(88, 64, [0.3, 0.5], seed=i, frac=0.05)
        y = synth.bernoulli_roll(88, 64, 0.2, seed=i)
        for t in (0.3, 0.5, 0.999999, 0.0):
            tp, fp, fn = of1.counts(p, y, 50, t)
            pred = (torch.from_numpy(p) > t).float().numpy()[:, :50].flatten()
            cm = confusion_matrix(y[:, :50].flatten(), pred, labels=[0, 1])
            assert (tp, fp, fn) == (cm[1, 1], cm[0, 1], cm[1, 0])
            assert of1.f1_from_counts(tp, fp, fn) == pytest.approx(
                f1_score(y[:, :50].flatten(), pred, zero_division=0), abs=1e-15)
    assert of1.f1_from_counts(0, 0, 0) == 0.0


def test_frontend_oracle_vs_torchaudio_fixture():
    g = np.load(os.path.join(GOLDEN, "frontend_torchaudio.npz"))
    y = synth.piano_chord(int(g["k"]), n_samples=int(g["n_samples"]))
    mine = ofe.logmel(y)
    assert mine.shape == (320, 1 + int(g["n_samples"]) // 512) and mine.dtype == np.float32
    np.testing.assert_array_equal(mine, g["logmel_oracle"])          # restatement is deterministic
    d = np.abs(mine - g["logmel_torchaudio"])
    assert d.max() < 2e-2 and d.mean() < 2e-4                       # fp32-FFT vs fp64-FFT noise only
    assert np.abs(mine - ofe.logmel_f64(y)).max() < 1e-4


def test_frontend_oracle_vs_transformers_audio_utils():
    """A second independent implementation of the librosa recipe: `transformers.audio_utils` (HuggingFace's numpy port of
    librosa.filters.mel / melspectrogram / power_to_db, installed in this image -- librosa itself is not).  Same slaney
    filterbank to 1e-8 and, with the arguments main.py:117-125 implies (periodic Hann 2048, hop 512, centred, constant
    padding, power 2, amin 1e-10, top_db 80), the same dB spectrogram to 1e-4 dB on chord, noise and silence."""
    au = pytest.importorskip("transformers.audio_utils")
    from music_transcription_b200 import synth
    from oracle import frontend as ofe
    for n_mels in (320, 229, 64):
        fb_hf = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=n_mels, min_frequency=0.0, max_frequency=8000.0,
                                   sampling_rate=16000, norm="slaney", mel_scale="slaney")
        assert np.abs(fb_hf.T - ofe.mel_filterbank(n_mels=n_mels)).max() < 1e-7
    fb_hf = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=320, min_frequency=0.0, max_frequency=8000.0,
                               sampling_rate=16000, norm="slaney", mel_scale="slaney")
    win = au.window_function(2048, "hann", periodic=True)
    rng = np.random.default_rng(3)
    for y in (synth.piano_chord(0, n_samples=64000), rng.uniform(-1, 1, 30000).astype(np.float32), np.zeros(9000, np.float32)):
        S = au.spectrogram(y.astype(np.float64), win, frame_length=2048, hop_length=512, fft_length=2048, power=2.0, center=True,
                           pad_mode="constant", onesided=True, mel_filters=fb_hf, mel_floor=1e-10, log_mel="dB", reference=1.0,
                           min_value=1e-10, db_range=80.0)
        ref = ofe.logmel(y)
        assert S.shape == ref.shape and np.abs(S - ref).max() < 1e-4, np.abs(S - ref).max()


def test_frontend_filterbank_properties():
    fb = ofe.mel_filterbank()
    assert fb.shape == (320, 1025) and fb.dtype == np.float32
    nnz = (fb > 0).sum(1)
    assert nnz.min() >= 2 and nnz.max() <= 20 and 0.005 < (fb > 0).mean() < 0.007
    assert abs(float(fb.max()) - 0.1058) < 1e-3
    # frame count and top_db floor
    y = synth.piano_chord(1, n_samples=480000)
    db = ofe.logmel(y)
    assert db.shape == (320, 938)
    assert db.min() == pytest.approx(db.max() - 80.0, abs=1e-4) or db.min() > db.max() - 80.0
    sil = ofe.logmel(np.zeros(48000, np.float32))
    assert np.allclose(sil, -100.0, atol=1e-4)        # amin floor, max-80 is below it


def _loss_cases():
    g = np.load(os.path.join(GOLDEN, "loss_reference.npz"))
    for name in g["names"]:
        name = str(name)
        heads = bool(g[f"{name}.heads"])
        logits = {k: torch.from_numpy(g[f"{name}.{k}"]) for k in ("frame", "onset", "offset")}
        lengths = g[f"{name}.lengths"]
        yield (name, logits if heads else logits["frame"], torch.from_numpy(g[f"{name}.roll"]),
               None if lengths.size == 0 else torch.from_numpy(lengths), float(g[f"{name}.loss"]))


def test_loss_oracle_matches_reference_compute_loss():
    from oracle import losses as olosses
    n = 0
    for name, logits, roll, lengths, ref in _loss_cases():
        got = float(olosses.compute_loss(logits, roll, lengths))
        assert abs(got - ref) <= 1e-6 * max(1.0, abs(ref)), (name, got, ref)
        n += 1
    assert n == 9
