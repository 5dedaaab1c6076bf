"""Audio loading: WAVE decode and the polyphase filter on the CPU; the GPU resampler
against scipy.signal.resample_poly (float64) under -m gpu."""
import numpy as np
import pytest
import torch

from music_transcription_b200 import audio


def test_polyphase_taps_equal_scipy_default_filter():
    from scipy.signal import firwin
    for up, down in ((160, 441), (1, 3), (2, 1), (320, 441)):
        mr = max(up, down)
        want = firwin(2 * 10 * mr + 1, 1.0 / mr, window=("kaiser", 5.0)) * up
        assert np.allclose(audio.polyphase_taps(up, down), want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("dtype,ch", [("int16", 1), ("int16", 2), ("int32", 2), ("float32", 1), ("uint8", 1)])
def test_wav_decode_matches_scipy_reader(tmp_path, dtype, ch):
    from scipy.io import wavfile
    rng = np.random.default_rng(1)
    n = 1000
    if dtype == "float32":
        x = rng.uniform(-1, 1, (n, ch)).astype(np.float32)
        want = x
    elif dtype == "uint8":
        x = rng.integers(0, 256, (n, ch)).astype(np.uint8)
        want = (x.astype(np.float32) - 128) / 128
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, (n, ch)).astype(dtype)
        want = (x.astype(np.float64) / (-float(info.min))).astype(np.float32)
    path = tmp_path / "a.wav"
    wavfile.write(path, 22050, x if ch > 1 else x[:, 0])
    got, sr = audio.load_wav(str(path))
    assert sr == 22050 and got.shape == (n, ch) and np.array_equal(got, want.reshape(n, ch))


def test_wav_decode_24_bit_and_rejects_other_containers(tmp_path):
    import struct
    vals = np.array([0, 1, -1, 8388607, -8388608, 123456, -654321], np.int32)
    raw = b"".join(struct.pack("<i", int(v))[:3] for v in vals)
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 48000, 144000, 3, 24)
    p = tmp_path / "b.wav"
    p.write_bytes(hdr + b"data" + struct.pack("<I", len(raw)) + raw + b"\0")
    got, sr = audio.load_wav(str(p))
    assert sr == 48000 and np.array_equal(got[:, 0], (vals / 8388608.0).astype(np.float32))
    q = tmp_path / "c.mp3"
    q.write_bytes(b"ID3\x03" + b"\0" * 64)
    with pytest.raises(ValueError):
        audio.load_wav(str(q))


@pytest.mark.gpu
@pytest.mark.parametrize("orig,n", [(44100, 44100 * 3 + 17), (48000, 100000), (22050, 50001), (8000, 12345), (16000, 999),
                                    (44100, 300)])
def test_gpu_resampler_matches_scipy_resample_poly(orig, n):
    from math import gcd
    from scipy.signal import resample_poly
    rng = np.random.default_rng(orig + n)
    t = np.arange(n) / orig
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3000 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    got = audio.resample(x, orig, 16000).cpu().numpy()
    if orig == 16000:
        assert np.array_equal(got, x)
        return
    g = gcd(orig, 16000)
    want = resample_poly(x.astype(np.float64), 16000 // g, orig // g)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5


@pytest.mark.gpu
def test_transcribe_audio_writes_a_midi_file(tmp_path):
    from scipy.io import wavfile
    from music_transcription_b200 import synth
    from music_transcription_b200.transcription_model import TranscriptionModel
    from oracle import smf as osmf
    sr_file = 44100
    t = np.arange(int(sr_file * 33.0)) / sr_file                      # 33 s -> two 30-s chunks, the second zero padded
    y = 0.3 * np.sin(2 * np.pi * 261.63 * t) * np.exp(-1.5 * (t % 2.0))
    stereo = np.stack([y, 0.5 * y], axis=1)
    wav_path = tmp_path / "take.wav"
    wavfile.write(wav_path, sr_file, (stereo * 32767).astype(np.int16))
    m = TranscriptionModel("cnn_rnn", n_mels=320, hidden_size=128, num_layers=1, device="cuda")
    m.load_state_dict(synth.synth_state_dict("cnn_rnn", 320, 128, 1, seed=2, gain=2.0))
    out = audio.transcribe_audio(wav_path, m)
    assert str(out).endswith("take_transcription.mid")
    parsed = osmf.parse(open(out, "rb").read())
    assert parsed["division"] == 220 and len(parsed["tracks"]) == 2
    notes = osmf.notes_from(parsed)
    assert all(21 <= p <= 108 and e > s for p, _, s, e in notes)
    assert max((e for *_, e in notes), default=0) <= round(2 * 938 / 31.25 * 440) + 1


def test_pcm16_reader_and_tensor_chunk_split(tmp_path):
    from scipy.io import wavfile
    from music_transcription_b200 import pipeline
    rng = np.random.default_rng(3)
    pcm = rng.integers(-32768, 32767, size=(1000, 2), dtype=np.int16)
    p = tmp_path / "s.wav"
    wavfile.write(p, 22050, pcm)
    raw, sr = audio.load_wav_pcm16(str(p))
    assert sr == 22050 and raw.dtype == np.int16 and np.array_equal(raw, pcm)
    x, _ = audio.load_wav(str(p))
    assert np.array_equal(x, pcm.astype(np.float32) / 32768.0)
    pf = tmp_path / "f.wav"
    wavfile.write(pf, 16000, rng.standard_normal(100).astype(np.float32))
    assert audio.load_wav_pcm16(str(pf)) is None                       # float WAVE: not the int16 fast path
    # a torch signal is chunked exactly like the reference's numpy loop (main.py:82-97)
    y = rng.standard_normal(2 * 480000 + 123).astype(np.float32)
    ref = np.stack(pipeline.split_audio_into_chunks(y))
    got = pipeline.split_audio_into_chunks(torch.from_numpy(y))
    assert got.shape == (3, 480000) and np.array_equal(got.numpy(), ref)
    assert pipeline.split_audio_into_chunks(torch.zeros(0)).shape == (0, 480000)


@pytest.mark.gpu
@pytest.mark.parametrize("ch", [1, 2, 3])
def test_pcm16_upload_path_is_bit_identical_to_host_decode(tmp_path, ch):
    """load_audio uploads 16-bit PCM as int16 and converts / mixes on the GPU (amt_pcm16_to_mono_f32): the result must
    equal the host float32 decode + numpy mean fed through the same resampler, bit for bit."""
    from scipy.io import wavfile
    rng = np.random.default_rng(ch)
    pcm = rng.integers(-32768, 32767, size=(50001, ch), dtype=np.int16)
    pcm[:4] = [[32767] * ch, [-32768] * ch, [0] * ch, [1] * ch]
    p = tmp_path / "s.wav"
    wavfile.write(p, 44100, pcm if ch > 1 else pcm[:, 0])
    got, sr = audio.load_audio(str(p), device="cuda")
    x, file_sr = audio.load_wav(str(p))
    y = x.mean(axis=1, dtype=np.float32) if ch > 1 else x[:, 0]
    want = audio.resample(y, file_sr, 16000, "cuda")
    assert sr == 16000 and got.shape == want.shape and torch.equal(got, want)
