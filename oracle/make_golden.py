"""ORACLE tooling: generate tests/golden/*.npz by RUNNING THE UNMODIFIED
REFERENCE in this container (it imports /root/reference, which does not exist
on the GPU box -- so only the committed vectors travel).

    python oracle/make_golden.py            # rewrites tests/golden/

What is executed from the reference:
  * models.transcription_model.TranscriptionModel (both CNN-RNN types), loaded
    with ``load_state_dict(strict=True)`` from music_transcription_b200.synth
    -> pins the checkpoint key set and the forward numerics of oracle/model.py.
  * main.pianoroll_to_midi, with stand-in ``librosa``/``pretty_midi`` modules
    (both are absent offline; the stand-in only records Note(...) calls)
    -> pins oracle/notes.py.
  * scripts/evaluate.py evaluate_at_threshold / run_threshold_tuning with the
    globals the script imports under ``__main__`` injected, and a stub model
    -> pins oracle/f1.py and the sweep schedule.
  * TranscriptionModel.compute_loss (models/transcription_model.py:110-217) on seeded logits / rolls
    -> pins oracle/losses.py.
  * torchaudio (independent implementation, not the reference) as a
    cross-check of oracle/frontend.py -- librosa itself cannot be run, so the
    frontend stays "parity unpinned".
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from music_transcription_b200 import synth  # noqa: E402

MODEL_CASES = [
    # name, model_type, n_mels, hidden, layers, B, T, attention, heads
    ("small_a", "cnn_rnn", 64, 128, 2, 2, 96, True, True),
    ("large_a", "cnn_rnn_large", 64, 128, 2, 2, 96, True, True),
    ("large_noattn", "cnn_rnn_large", 64, 128, 1, 1, 50, False, True),
    ("large_nohead", "cnn_rnn_large", 64, 128, 2, 1, 50, True, False),
    ("large_oddmel", "cnn_rnn_large", 37, 128, 2, 1, 41, True, True),
    ("small_oddmel", "cnn_rnn", 37, 128, 3, 1, 41, True, True),
]


def _ref_modules():
    sys.path.insert(0, REF)
    from models.transcription_model import TranscriptionModel  # type: ignore
    return TranscriptionModel


def gen_models():
    TM = _ref_modules()
    for name, mt, n_mels, H, L, B, T, attn, heads in MODEL_CASES:
        sd = synth.synth_state_dict(mt, n_mels, H, L, seed=3, use_attention=attn,
                                    use_onset_offset_heads=heads)
        m = TM(model_type=mt, n_mels=n_mels, hidden_size=H, num_layers=L, dropout=0.2, device="cpu",
               use_attention=attn, use_onset_offset_heads=heads)
        ref_keys = list(m.state_dict().keys())
        assert ref_keys == list(sd.keys()), (name, set(ref_keys) ^ set(sd.keys()))
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = synth.synth_logmel(B, n_mels, T, seed=11)
        with torch.no_grad():
            out = {"frame": m(x).numpy()}
            if mt.endswith("large") and heads:
                all_h = m(x, return_all_heads=True)
                assert np.array_equal(all_h["frame"].numpy(), out["frame"])
                out["onset"] = all_h["onset"].numpy()
                out["offset"] = all_h["offset"].numpy()
            pred = m.predict(x, threshold=0.5).numpy()
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"),
                            cfg=np.array([n_mels, H, L, B, T, int(attn), int(heads), 3, 11]),
                            model_type=mt, keys=np.array(ref_keys), x=x.numpy(), pred=pred, **out)
        print("model", name, out["frame"].shape, float(np.abs(out["frame"]).max()))
    # T == 0: the zero-length guard (reference cnn_rnn_model.py:65-66,297-304) is
    # unreachable -- conv2d rejects a (n_mels x 0) input first.  Record that.
    m = TM(model_type="cnn_rnn_large", n_mels=64, hidden_size=128, num_layers=1, device="cpu").eval()
    try:
        with torch.no_grad():
            m(torch.zeros(2, 1, 64, 0))
        raised = ""
    except RuntimeError as e:
        raised = str(e)[:120]
    print("T=0 ->", raised or "no error")
    np.savez_compressed(os.path.join(OUT, "model_T0.npz"), raised=np.array(raised))


CANON = dict(n_mels=320, hidden=512, layers=3, T=938, seed=1, gain=3 ** -0.5)


def canon_input(ks):
    """The canonical-shape fixture input: log-mel of the SURVEY 8(d) chord chunks (oracle frontend), rounded to the
    float16 grid so the fixture stores it exactly in half the bytes.  (B,1,320,938) float32."""
    from oracle import frontend as fe
    x = np.stack([fe.logmel(synth.piano_chord(int(k))) for k in ks])[:, None]
    return torch.from_numpy(x.astype(np.float16).astype(np.float32))


def gen_canonical():
    """Reference-run fixtures AT THE CANONICAL SHAPES (main.py:16-24: n_mels 320, hidden 512, 3 layers, T 938), so the
    16-CTA-cluster LSTM (H = 512 / 256), the head-dim-192 attention and the K = 10240 projection are pinned by the real
    reference modules, not only by the port: CNNRNNModelLarge on 2 chord chunks (three heads) and CNNRNNModel (36 M) on 1."""
    TM = _ref_modules()
    c = CANON
    for name, mt, ks in (("large", "cnn_rnn_large", [0, 3]), ("small", "cnn_rnn", [0])):
        sd = synth.synth_state_dict(mt, c["n_mels"], c["hidden"], c["layers"], seed=c["seed"], gain=c["gain"])
        m = TM(model_type=mt, n_mels=c["n_mels"], hidden_size=c["hidden"], num_layers=c["layers"], dropout=0.2, device="cpu")
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = canon_input(ks)
        with torch.no_grad():
            out = {"frame": m(x).numpy()}
            if mt.endswith("large"):
                allh = m(x, return_all_heads=True)
                assert np.array_equal(allh["frame"].numpy(), out["frame"])
                out["onset"], out["offset"] = allh["onset"].numpy(), allh["offset"].numpy()
        np.savez_compressed(os.path.join(OUT, f"canon_{name}.npz"), model_type=mt, chunks=np.array(ks),
                            cfg=np.array([c["n_mels"], c["hidden"], c["layers"], len(ks), c["T"], c["seed"]]),
                            gain=np.float64(c["gain"]), x=x.numpy().astype(np.float16), **out)
        print("canonical", name, out["frame"].shape, "logit std", float(out["frame"].std()))


class _FakeNote:
    def __init__(self, velocity, pitch, start, end):
        self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end


class _FakeInstrument:
    def __init__(self, program=0):
        self.program, self.notes = program, []


class _FakeMidi:
    def __init__(self):
        self.instruments = []


def _import_ref_main():
    pm = types.ModuleType("pretty_midi")
    pm.PrettyMIDI, pm.Instrument, pm.Note = _FakeMidi, _FakeInstrument, _FakeNote
    sys.modules.setdefault("pretty_midi", pm)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    spec = importlib.util.spec_from_file_location("ref_main", os.path.join(REF, "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_notes():
    ref_main = _import_ref_main()
    cases = {}
    rng = np.random.default_rng(5)
    rolls = {
        "random": (rng.random((88, 300)) < 0.3).astype(np.float32),
        "sparse": (rng.random((88, 938)) < 0.02).astype(np.float32),
        "full": np.ones((88, 40), np.float32),
        "empty": np.zeros((88, 40), np.float32),
        "edges": np.zeros((88, 64), np.float32),
    }
    rolls["edges"][0, 0] = 1
    rolls["edges"][1, 63] = 1
    rolls["edges"][2, :] = 1
    rolls["edges"][3, 10:20] = 1
    rolls["edges"][87, 62:] = 1
    # two-chunk seam case through the reference's own combine_piano_rolls
    a = (rng.random((88, 938)) < 0.1).astype(np.float32)
    b = (rng.random((88, 938)) < 0.1).astype(np.float32)
    a[5, 930:] = 1
    b[5, :7] = 1
    rolls["seam"] = ref_main.combine_piano_rolls([a, b])
    assert ref_main.combine_piano_rolls([a]) is a
    fs = 16000 / 512
    for name, roll in rolls.items():
        midi = ref_main.pianoroll_to_midi(roll, fs, min_midi=21)
        notes = midi.instruments[0].notes
        cases[name + "_roll"] = roll.astype(np.uint8)
        cases[name + "_pitch"] = np.array([n.pitch for n in notes], dtype=np.int32)
        cases[name + "_start"] = np.array([n.start for n in notes], dtype=np.float64)
        cases[name + "_end"] = np.array([n.end for n in notes], dtype=np.float64)
        cases[name + "_vel"] = np.array([n.velocity for n in notes], dtype=np.int32)
        print("notes", name, len(notes))
    cases["seam_a"] = a.astype(np.uint8)
    cases["seam_b"] = b.astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "notes_reference.npz"), **cases)


class _StubModel(torch.nn.Module):
    """Returns pre-baked logits in call order (the reference re-runs the model
    for every threshold, scripts/evaluate.py:538)."""

    def __init__(self, logits_list):
        super().__init__()
        self.logits_list, self.i = logits_list, 0

    def forward(self, mel):
        out = self.logits_list[self.i % len(self.logits_list)]
        self.i += 1
        return out


def gen_f1():
    from sklearn.metrics import f1_score
    spec = importlib.util.spec_from_file_location("ref_eval", os.path.join(REF, "scripts", "evaluate.py"))
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)

    class _Tq:
        def __call__(self, it, **kw):
            return it

        @staticmethod
        def write(s):
            pass
    ev.np, ev.torch, ev.f1_score, ev.tqdm = np, torch, f1_score, _Tq()

    n_pieces, Tmax = 6, 120
    lengths = [120, 117, 60, 120, 1, 90]
    thr_grid = np.linspace(0.01, 0.99, 100)
    probs, rolls, batches, logits = [], [], [], []
    for i in range(n_pieces):
        p = synth.planted_probs(88, Tmax, thr_grid, seed=i, frac=0.02)
        p = np.clip(p, 1e-6, 1 - 1e-6).astype(np.float32)
        if i == 3:
            p[:] = 0.001            # nothing predicted -> F1 0 by zero_division
        y = synth.bernoulli_roll(88, Tmax, 0.05, seed=i)
        if i == 4:
            y[:] = 0
        lg = torch.logit(torch.from_numpy(p).double()).float()
        # the oracle / CUDA kernels take probabilities: store sigmoid(logits) as torch computes it
        p_eff = torch.sigmoid(lg).numpy()
        probs.append(p_eff)
        rolls.append(y)
        logits.append(lg[None])
        batches.append((torch.zeros(1, 1, 8, Tmax), torch.from_numpy(y)[None], torch.tensor([lengths[i]])))

    info = {"device": "cpu"}
    f1_at = []
    for t in [0.5, 0.35000000000000003, 0.1, float(np.float32(0.1)), 0.9500000000000002]:
        f1_at.append((t, ev.evaluate_at_threshold(_StubModel(logits), batches, info, t)))
    args = types.SimpleNamespace(tune_range=[0.05, 0.95], tune_step=0.1, tune_min_step=0.01, tune_rounds=6,
                                 model="stub", split="test", subset=None)
    info2 = {"device": "cpu", "data_source": "synthetic"}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        best_t, best_f1 = ev.run_threshold_tuning(args, _StubModel(logits), batches, info2)
    print("f1 sweep", best_t, best_f1, f1_at)
    np.savez_compressed(os.path.join(OUT, "f1_reference.npz"),
                        probs=np.stack(probs), rolls=np.stack(rolls).astype(np.uint8), lengths=np.array(lengths),
                        at_t=np.array([a for a, _ in f1_at], dtype=np.float64),
                        at_f1=np.array([b for _, b in f1_at], dtype=np.float64),
                        best_t=np.float64(best_t), best_f1=np.float64(best_f1))


def gen_frontend():
    import torchaudio
    sys.path.insert(0, ROOT)
    from oracle import frontend as fe
    y = synth.piano_chord(0, n_samples=64000)
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=320,
                                              center=True, pad_mode="constant", power=2.0, norm="slaney",
                                              mel_scale="slaney", f_min=0.0, f_max=8000.0)
    db = torchaudio.transforms.AmplitudeToDB("power", top_db=80.0)
    ta = db(ms(torch.from_numpy(y))[None])[0].numpy()
    mine = fe.logmel(y)
    fb_ta = ms.mel_scale.fb.numpy().T
    print("frontend vs torchaudio: max", np.abs(ta - mine).max(), "mean", np.abs(ta - mine).mean(),
          "fb", np.abs(fb_ta - fe.mel_filterbank()).max())
    np.savez_compressed(os.path.join(OUT, "frontend_torchaudio.npz"), n_samples=64000, k=0,
                        logmel_torchaudio=ta.astype(np.float32), logmel_oracle=mine,
                        fb_rowsum=fe.mel_filterbank().sum(1), fb_nnz=(fe.mel_filterbank() > 0).sum(1))


# (name, B, T_logits, T_targets, three heads, lengths or None)
LOSS_CASES = [
    ("single", 3, 50, 50, False, None),
    ("single_masked", 4, 61, 61, False, [61, 17, 0, 40]),
    ("single_interp", 2, 47, 94, False, [94, 30]),
    ("single_interp_down", 2, 100, 33, False, None),
    ("heads", 3, 50, 50, True, None),
    ("heads_masked", 4, 61, 61, True, [61, 17, 1, 200]),
    ("heads_interp", 2, 30, 45, True, [45, 11]),
    ("heads_T1", 2, 1, 1, True, None),
    ("all_masked", 2, 20, 20, True, [0, 0]),
]


def gen_losses():
    """TranscriptionModel.compute_loss of the reference itself (models/transcription_model.py:110-217) on seeded
    logits / rolls -> pins oracle/losses.py and the amt_bce_loss kernel."""
    TM = _ref_modules()
    small = TM(model_type="cnn_rnn", n_mels=64, hidden_size=128, num_layers=1, device="cpu")
    out = {}
    for i, (name, B, Tl, Tt, heads, lengths) in enumerate(LOSS_CASES):
        g = torch.Generator().manual_seed(100 + i)
        logits = {k: torch.randn(B, 88, Tl, generator=g) * 3.0 for k in ("frame", "onset", "offset")}
        # rolls with runs (notes), so onset / offset targets are non-trivial
        roll = (torch.rand(B, 88, Tt, generator=g) < 0.3).float()
        roll[:, :, 1:] = torch.maximum(roll[:, :, 1:], (roll[:, :, :-1] * (torch.rand(B, 88, Tt - 1, generator=g) < 0.6)).float()) if Tt > 1 else roll[:, :, 1:]
        lt = None if lengths is None else torch.tensor(lengths)
        with torch.no_grad():
            val = small.compute_loss(logits if heads else logits["frame"], roll, lt)
        out[f"{name}.frame"], out[f"{name}.onset"], out[f"{name}.offset"] = (logits[k].numpy() for k in ("frame", "onset", "offset"))
        out[f"{name}.roll"] = roll.numpy()
        out[f"{name}.lengths"] = np.array([] if lengths is None else lengths, dtype=np.int64)
        out[f"{name}.heads"] = np.array(int(heads))
        out[f"{name}.loss"] = np.array(float(val), dtype=np.float64)
        print("loss", name, float(val))
    np.savez_compressed(os.path.join(OUT, "loss_reference.npz"), names=np.array([c[0] for c in LOSS_CASES]), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == "losses":
        gen_losses()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "canonical":
        torch.set_num_threads(os.cpu_count() or 4)
        gen_canonical()
        sys.exit(0)
    gen_models()
    gen_canonical()
    gen_losses()
    gen_notes()
    gen_f1()
    gen_frontend()
