"""CPU: every function / method the drop-in mirrors has the reference's signature -- the reference's parameters, in
order, with the same names and defaults (our extra parameters, if any, come after them and are optional), so that
`import main` -> `from music_transcription_b200 import main` really is a one-line change.
Needs /root/reference (present in the build container, absent on the GPU box -> skipped there)."""
import importlib.util
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _ref_main():
    for name in ("pretty_midi", "librosa"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_main_sig", os.path.join(REF, "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_evaluate():
    spec = importlib.util.spec_from_file_location("ref_eval_sig", os.path.join(REF, "scripts", "evaluate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _assert_compatible(ours, ref, what):
    po = [p for p in inspect.signature(ours).parameters.values() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
    pr = [p for p in inspect.signature(ref).parameters.values() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
    assert len(po) >= len(pr), (what, [p.name for p in po], [p.name for p in pr])
    for a, b in zip(po, pr):
        assert a.name == b.name, (what, a.name, b.name)
        assert a.default == b.default or (a.default is inspect._empty) == (b.default is inspect._empty) and a.default == b.default, \
            (what, a.name, a.default, b.default)
    for extra in po[len(pr):]:
        assert extra.default is not inspect._empty, (what, "extra parameter without a default", extra.name)
    ref_kw = any(p.kind == p.VAR_KEYWORD for p in inspect.signature(ref).parameters.values())
    our_kw = any(p.kind == p.VAR_KEYWORD for p in inspect.signature(ours).parameters.values())
    assert our_kw or not ref_kw, (what, "the reference accepts **kwargs")


def test_main_module_mirrors_reference_main():
    from music_transcription_b200 import main as ours
    ref = _ref_main()
    for fn in ("load_model", "split_audio_into_chunks", "audio_to_mel", "predict_chunk", "combine_piano_rolls", "pianoroll_to_midi",
               "transcribe_audio", "main"):
        _assert_compatible(getattr(ours, fn), getattr(ref, fn), f"main.{fn}")
    for const in ("MODEL_TYPE", "N_MELS", "HIDDEN_SIZE", "NUM_LAYERS", "DROPOUT", "SR", "HOP_LENGTH", "CHUNK_LENGTH", "THRESHOLD"):
        assert getattr(ours, const) == getattr(ref, const), const


def test_transcription_model_mirrors_reference_class():
    from music_transcription_b200.transcription_model import TranscriptionModel as Ours
    _ref_main()
    from models.transcription_model import TranscriptionModel as Ref  # type: ignore
    for meth in ("__init__", "forward", "predict", "compute_loss"):
        _assert_compatible(getattr(Ours, meth), getattr(Ref, meth), f"TranscriptionModel.{meth}")


def test_evaluate_mirrors_reference_script_functions():
    from music_transcription_b200 import evaluate as ours
    ref = _ref_evaluate()
    for fn in ("evaluate_at_threshold", "run_threshold_tuning"):
        _assert_compatible(getattr(ours, fn), getattr(ref, fn), f"evaluate.{fn}")
    from music_transcription_b200 import pipeline
    _assert_compatible(pipeline.pianoroll_to_midi, ref.pianoroll_to_midi, "evaluate.pianoroll_to_midi")


def test_unsupported_configurations_fail_at_construction():
    from music_transcription_b200.transcription_model import TranscriptionModel
    for kw in (dict(hidden_size=200), dict(hidden_size=768), dict(model_type="cnn_rnn_large", hidden_size=640), dict(num_layers=0)):
        with pytest.raises(ValueError):
            TranscriptionModel(**{"model_type": "cnn_rnn", "n_mels": 64, "device": "cpu", **kw})
    TranscriptionModel("cnn_rnn_large", n_mels=64, hidden_size=640, use_attention=False, device="cpu")      # 640 is fine without attention
