"""CPU: the C-ABI library builds/loads without a GPU and exports every symbol include/amt.h
declares; argument validation that needs no device works; compute entry points fail loudly
(no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import os
import re

import pytest
import torch

from music_transcription_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "amt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/amt.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert lib.amt_version().decode().startswith("amt-sm100")


def test_argument_errors_are_reported_not_thrown():
    lib = _lib.lib()
    cfg = _lib.ModelConfig(7, 320, 512, 3, 8, 1, 1)
    h = C.c_void_p()
    assert lib.amt_model_create(C.byref(cfg), C.byref(h)) == _lib.AMT_ERR_ARG
    assert b"kind" in lib.amt_last_error()
    cfg = _lib.ModelConfig(1, 320, 500, 3, 8, 1, 1)
    with pytest.raises(ValueError):
        _lib.check(lib.amt_model_create(C.byref(cfg), C.byref(h)))
    cfg = _lib.ModelConfig(1, 320, 512, 3, 8, 1, 1)
    _lib.check(lib.amt_model_create(C.byref(cfg), C.byref(h)))
    assert lib.amt_model_finalize(h) == _lib.AMT_ERR_STATE and b"missing tensor" in lib.amt_last_error()
    ws = lib.amt_model_workspace_bytes(h, 16, 938)
    assert 2 * 2 ** 30 < ws < 8 * 2 ** 30           # ~200 MB of activations per 30-s chunk
    lib.amt_model_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback_without_a_gpu():
    lib = _lib.lib()
    assert lib.amt_device_check() == _lib.AMT_ERR_DEVICE
    x = torch.zeros(128, 64, dtype=torch.bfloat16)
    st = lib.amt_gemm_bf16(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 128, 64, 64, 64, 0, 0, None)
    assert st == _lib.AMT_ERR_DEVICE
    from music_transcription_b200.transcription_model import TranscriptionModel
    m = TranscriptionModel("cnn_rnn", n_mels=64, hidden_size=128, num_layers=1, device="cpu")
    with pytest.raises(_lib.AmtError):
        m(torch.zeros(1, 1, 64, 10))


def test_drop_in_state_dict_keys_match_reference_spec():
    from music_transcription_b200 import synth
    from music_transcription_b200.transcription_model import TranscriptionModel
    for mt, attn, heads in (("cnn_rnn", True, True), ("cnn_rnn_large", True, True), ("large", False, True),
                            ("cnn_rnn_large", True, False)):
        m = TranscriptionModel(mt, n_mels=64, hidden_size=128, num_layers=2, device="cpu", use_attention=attn,
                               use_onset_offset_heads=heads)
        spec = synth.state_dict_spec(mt, 64, 128, 2, attn, heads)
        sd = m.state_dict()
        assert list(sd.keys()) == [k for k, _, _ in spec]
        for k, shape, _ in spec:
            assert tuple(sd[k].shape) == tuple(shape), k
        m.load_state_dict(synth.synth_state_dict(mt, 64, 128, 2, 1, attn, heads), strict=True)
        assert m.model_type == mt and m.device == "cpu" and m.use_onset_offset_heads == heads
    with pytest.raises(ValueError):
        TranscriptionModel("bogus")


def test_product_package_never_touches_the_oracle_or_the_reference_tree():
    """oracle/ is test infrastructure: nothing under music_transcription_b200/ (Python or CUDA) may import it or
    read /root/reference; the only other allowed users are bench.py's baseline legs and __graft_entry__.smoke()."""
    pkg = os.path.join(ROOT, "music_transcription_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        if "build" in os.path.relpath(d, pkg).split(os.sep):
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            src = open(os.path.join(d, f), errors="replace").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "/root/reference" in src:
                offenders.append(os.path.relpath(os.path.join(d, f), ROOT))
    assert offenders == []
    for f in os.listdir(os.path.join(ROOT, "tools")):         # helper scripts are not a back door either
        if f.endswith((".py", ".sh")) and not f.startswith("_"):
            src = open(os.path.join(ROOT, "tools", f), errors="replace").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) and "/root/reference" not in src, f
    for f in ("bench.py", "__graft_entry__.py"):          # nothing the GPU box runs may read the reference tree
        assert "/root/reference" not in open(os.path.join(ROOT, f)).read()


def test_deferred_logmel_host_logic_and_argument_checks():
    """pipeline.DeferredLogMel (unfloored dB + per-chunk maxima): floored() is power_to_db's max - top_db clamp
    (reference main.py:125) per chunk; the new C entry points refuse bad arguments without touching a GPU."""
    import ctypes as C

    import torch
    from music_transcription_b200 import _lib, pipeline
    mel = torch.tensor([[[[0.0, -50.0], [-90.0, -100.0]]], [[[-10.0, -95.0], [-200.0, -89.0]]]])      # (2, 1, 2, 2)
    d = pipeline.DeferredLogMel(mel, torch.tensor([0.0, -10.0]), 80.0)
    assert d.shape == mel.shape and d.device == mel.device
    want = torch.tensor([[[[0.0, -50.0], [-80.0, -80.0]]], [[[-10.0, -90.0], [-90.0, -89.0]]]])
    assert torch.equal(d.floored(), want)
    L = _lib.lib()
    assert L.amt_model_forward_db(None, None, None, 80.0, 1, 1, None, None, None, None, 0, None) == _lib.AMT_ERR_ARG
    assert L.amt_onset_notes(None, None, None, 1, 88, 10, None, 0, None, None, 0, None) == _lib.AMT_ERR_ARG
    assert L.amt_onset_notes_scratch_ints(88) == 88 and L.amt_onset_notes_scratch_ints(0) == 0
    buf = (C.c_int32 * 4)()
    st = L.amt_onset_notes(buf, buf, None, 1, 2000, 10, buf, 0, buf, buf, 4, None)                  # too many pitches
    assert st == _lib.AMT_ERR_ARG and b"1024" in L.amt_last_error()
