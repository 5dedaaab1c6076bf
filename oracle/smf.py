"""Test infrastructure only (never imported by the product path): an independent Standard MIDI File
READER used to check ``music_transcription_b200.smf`` by round trip -- it parses header, tracks,
variable-length deltas, meta events and channel messages (with and without running status) and returns
absolute-tick event lists, from which notes are re-paired the way ``pretty_midi`` loads them
(a note_on with velocity 0 closes the oldest open note of that pitch)."""
import struct


def _read_varlen(b, i):
    v = 0
    while True:
        c = b[i]
        i += 1
        v = (v << 7) | (c & 0x7F)
        if not c & 0x80:
            return v, i


def parse(data: bytes):
    assert data[:4] == b"MThd"
    hlen, fmt, ntrk, div = struct.unpack(">IHHH", data[4:14])
    assert hlen == 6
    i = 14
    tracks = []
    for _ in range(ntrk):
        assert data[i:i + 4] == b"MTrk"
        n = struct.unpack(">I", data[i + 4:i + 8])[0]
        body = data[i + 8:i + 8 + n]
        i += 8 + n
        j, now, status, ev = 0, 0, None, []
        while j < len(body):
            d, j = _read_varlen(body, j)
            now += d
            c = body[j]
            if c == 0xFF:
                kind = body[j + 1]
                ln, k = _read_varlen(body, j + 2)
                ev.append((now, "meta", kind, bytes(body[k:k + ln])))
                j = k + ln
                continue
            if c & 0x80:
                status = c
                j += 1
            hi = status & 0xF0
            nbytes = 1 if hi in (0xC0, 0xD0) else 2
            ev.append((now, "msg", status, bytes(body[j:j + nbytes])))
            j += nbytes
        tracks.append(ev)
    assert i == len(data)
    return {"format": fmt, "division": div, "tracks": tracks}


def notes_from(parsed, track=1):
    """(pitch, velocity, start_tick, end_tick) in order of note end, like pretty_midi's loader."""
    open_notes, out = {}, []
    for tick, kind, status, payload in parsed["tracks"][track]:
        if kind != "msg" or status & 0xF0 not in (0x90, 0x80):
            continue
        pitch, vel = payload
        if status & 0xF0 == 0x90 and vel > 0:
            open_notes.setdefault(pitch, []).append((tick, vel))
        elif open_notes.get(pitch):
            s, v = open_notes[pitch].pop(0)
            out.append((pitch, v, s, tick))
    return out
