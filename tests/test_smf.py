"""SMF writer: hand-assembled golden bytes + round trip through an independent parser."""
import numpy as np

from music_transcription_b200 import pipeline, smf
from oracle import notes as onotes, smf as osmf


def _notelist(roll, fs):
    """What pipeline.pianoroll_to_midi returns, with the grouping done by the oracle (no GPU in this suite)."""
    return pipeline.NoteList(onotes.group_notes(roll), fs)


def test_time_to_tick_is_round_half_even_of_440_ticks_per_second():
    fs = 16000 / 512
    assert smf.time_to_tick(0.0) == 0
    assert smf.time_to_tick(1.0) == 440
    assert smf.time_to_tick(32 / fs) == round(32 / fs * 440)
    assert smf.time_to_tick(0.5 / 440) == 0 and smf.time_to_tick(1.5 / 440) == 2      # ties to even


def test_golden_bytes_of_a_two_note_file():
    # C4 from frame 0 to 10 and the same pitch re-attacked at frame 10 (release must precede the attack)
    roll = np.zeros((88, 40), np.float32)
    roll[39, 0:10] = 1
    roll[39, 11:20] = 1
    roll[43, 5:15] = 1
    nl = _notelist(roll, 31.25)
    got = smf.smf_bytes(nl.instruments[0].notes)
    t = lambda fr: round(fr / 31.25 * 440)
    want = bytearray(b"MThd\x00\x00\x00\x06\x00\x01\x00\x02\x00\xdc")
    trk0 = b"\x00\xff\x51\x03\x07\xa1\x20" + b"\x00\xff\x58\x04\x04\x02\x18\x08" + b"\x01\xff\x2f\x00"
    want += b"MTrk" + len(trk0).to_bytes(4, "big") + trk0
    ev = [(0, bytes([0xC0, 0])), (t(0), bytes([0x90, 60, 100])), (t(5), bytes([0x90, 64, 100])),
          (t(10), bytes([0x90, 60, 0])), (t(11), bytes([0x90, 60, 100])), (t(15), bytes([0x90, 64, 0])),
          (t(20), bytes([0x90, 60, 0]))]
    body, now = bytearray(), 0
    for tick, msg in ev:
        body += smf._varlen(tick - now) + msg
        now = tick
    body += b"\x01\xff\x2f\x00"
    want += b"MTrk" + len(body).to_bytes(4, "big") + body
    assert got == bytes(want)
    assert smf._varlen(0) == b"\x00" and smf._varlen(127) == b"\x7f" and smf._varlen(128) == b"\x81\x00"
    assert smf._varlen(0x3FFF) == b"\xff\x7f" and smf._varlen(0x4000) == b"\x81\x80\x00"


def test_round_trip_through_independent_parser(tmp_path):
    rng = np.random.default_rng(0)
    roll = (rng.random((88, 2000)) > 0.7).astype(np.float32)
    nl = _notelist(roll, 16000 / 512)
    path = tmp_path / "out.mid"
    nl.write(str(path))
    parsed = osmf.parse(path.read_bytes())
    assert parsed["format"] == 1 and parsed["division"] == 220 and len(parsed["tracks"]) == 2
    assert [e[2] for e in parsed["tracks"][0]] == [0x51, 0x58, 0x2F]
    back = sorted(osmf.notes_from(parsed))
    want = sorted((n.pitch, 100, smf.time_to_tick(n.start), smf.time_to_tick(n.end)) for n in nl.instruments[0].notes)
    assert back == want
    ticks = [e[0] for e in parsed["tracks"][1]]
    assert ticks == sorted(ticks) and parsed["tracks"][1][-1][2] == 0x2F and ticks[-1] == ticks[-2] + 1


def test_same_tick_ordering_release_before_attack_and_low_pitch_first():
    class N:
        def __init__(s, p, a, b): s.pitch, s.velocity, s.start, s.end = p, 100, a, b
    data = smf.smf_bytes([N(70, 0.0, 1.0), N(60, 1.0, 2.0), N(70, 1.0, 2.0)])
    ev = [e for e in osmf.parse(data)["tracks"][1] if e[1] == "msg" and e[0] == 440]
    assert [(e[3][0], e[3][1]) for e in ev] == [(60, 100), (70, 0), (70, 100)]
