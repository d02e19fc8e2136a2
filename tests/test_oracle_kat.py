"""Known-answer and self-consistency pins of the CPU oracle.

The reference ships no golden vectors and cannot run here (no JS engine), so these are the pins the repo creates
itself: header bytes and frame sizes derived from the bitstream syntax, the hand-traced per-call frame schedule of
Mp3Encoder, the silent-frame layout, and a full parse of the oracle's own bytes back to its quantised lines."""
import numpy as np
import pytest

import mp3_parse
from synth import make_signal


def test_c1_silence_mono_schedule(oracle):
    """Config C1: mono 44.1k/128k, 1 s of silence in 1152-sample encodeBuffer calls (README.md:72-79 chunking).
    Hand trace of Lame.js:202,1409-1414,1592-1663: call 1 -> 0 bytes, calls 2..39 -> one frame each, flush -> 2."""
    x = np.zeros(44100, dtype=np.int16)
    data, sizes, _ = oracle.encode_stream(1, 44100, 128, x, None, chunk=1152)
    assert sizes[0] == 0
    assert all(s in (417, 418) for s in sizes[1:39])
    assert sizes[39] == 418 + 418            # flush: two frames
    assert len(sizes) == 40
    # padding recurrence frac_SpF = 42300 (Encoder.js:442-446): first frame unpadded, then slot_lag pattern
    lag, frames = 42300, []
    for _ in range(40):
        lag -= 42300
        pad = 0
        if lag < 0:
            lag += 44100
            pad = 1
        frames.append(417 + pad)
    assert sum(frames) == len(data)
    off = 0
    for k, fl in enumerate(frames):
        f = data[off:off + fl]
        assert f[0] == 0xFF and f[1] == 0xFB, k
        assert f[2] == (0x90 | (2 if fl == 418 else 0)) and f[3] == 0xC4    # 128k, 44.1k, padding, mono
        # silent frame: 21 bytes header+side info, then "LAME" + version pushed through >> (03 00 09 08 00 04), zeros
        assert f[21:31] == bytes([0x4C, 0x41, 0x4D, 0x45, 3, 0, 9, 8, 0, 4]), k
        assert not any(f[31:]), k
        off += fl


@pytest.mark.parametrize("ch,sr,kbps,b2,b3,flen", [(2, 44100, 128, 0x90, 0x04, 417), (2, 48000, 320, 0xE4, 0x04, 960),
                                                   (1, 32000, 160, 0xA8, 0xC4, 720), (2, 44100, 123, 0x90, 0x04, 417)])
def test_header_bytes_and_frame_sizes(oracle, ch, sr, kbps, b2, b3, flen):
    l, r = make_signal("noise", 6000, sr, 5)
    data, _, _ = oracle.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)
    assert data[0] == 0xFF and data[1] == 0xFB and data[2] == b2 and data[3] == b3
    # every frame starts at the closed-form offset
    off, k = 0, 0
    while off < len(data):
        assert data[off] == 0xFF and (data[off + 1] & 0xF0) == 0xF0
        pad = (data[off + 2] >> 1) & 1
        off += flen + pad
        k += 1
    assert off == len(data)


@pytest.mark.parametrize("ch,sr,kbps,out_sr,b1,b2,flen", [
    (2, 22050, 64, 22050, 0xF3, 0x80, 208),     # MPEG-2: version bit 0, bitrate index 8 (64k), rate index 0; 72000*64/22050 = 208.98
    (1, 16000, 32, 16000, 0xF3, 0x48, 144),     # MPEG-2 16 kHz, bitrate index 4 (32k), rate index 2
    (1, 8000, 32, 8000, 0xE3, 0x48, 288),       # MPEG-2.5: sync 0xFFE, 72000*32/8000 = 288
    (2, 44100, 64, 24000, 0xF3, 0x84, 192),     # lamejs resamples 44.1k -> 24k at this bitrate (optimum_samplefreq)
])
def test_lsf_header_bytes_and_frame_sizes(oracle, ch, sr, kbps, out_sr, b1, b2, flen):
    """MPEG-2 / 2.5 KATs derived from the bitstream syntax (BitStream.js:259-282, :83-98): sync + version + layer +
    no-CRC byte, bitrate/samplerate index byte (padding bit masked), frame length floor(72000 * kbps / out_sr) (+1)."""
    assert oracle.out_samplerate(ch, sr, kbps) == out_sr
    l, r = make_signal("noise", 8 * 1152, sr, 5)
    data, _, _ = oracle.encode_stream(ch, sr, kbps, l, r if ch == 2 else None, chunk=1152)
    off, k = 0, 0
    while off < len(data):
        assert data[off] == 0xFF and data[off + 1] == b1 and (data[off + 2] & 0xFD) == b2, (k, data[off:off + 4].hex())
        off += flen + ((data[off + 2] >> 1) & 1)
        k += 1
    assert off == len(data) and k >= 8


@pytest.mark.parametrize("kind,ch,sr,kbps", [("noise", 2, 44100, 128), ("burst", 2, 44100, 128), ("white", 2, 48000, 320),
                                             ("sine", 1, 44100, 128), ("octave", 1, 44100, 128), ("sweep", 2, 44100, 128),
                                             ("white", 1, 44100, 320), ("noise", 2, 32000, 192)])
def test_bitstream_parses_back_to_l3enc(oracle, books, kind, ch, sr, kbps):
    """Decode the oracle's own bytes (independent parser) and compare with the encoder's internal state."""
    l, r = make_signal(kind, 40 * 1152 + 123, sr, 11)
    data, _, tr = oracle.encode_stream(ch, sr, kbps, l, r if ch == 2 else None, trace_frames=60)
    frames = mp3_parse.parse_stream(data, books)
    assert len(frames) == len(tr)
    seen = set()
    for f, t in zip(frames, tr):
        assert f["frame_len"] == t["frame_bytes"] and f["padding"] == t["padding"]
        for gr in range(2):
            for c in range(ch):
                g = f["gi"][gr][c]
                seen.add(int(t["blocktype"][gr][c]))
                assert g["block_type"] == t["blocktype"][gr][c]
                assert g["global_gain"] == t["global_gain"][gr][c]
                assert g["part2_length"] == t["part2_length"][gr][c]
                assert g["part2_3_length"] == t["part2_3_length"][gr][c] + t["part2_length"][gr][c]
                assert g["part2_3_length"] <= 4095
                assert np.array_equal(np.abs(g["ix"]), t["l3_enc"][gr][c])
                # signs follow the MDCT spectrum (quantiser reorder only affects short blocks)
                if g["block_type"] != 2:
                    nz = g["ix"] != 0
                    assert np.array_equal(g["ix"][nz] < 0, t["xr"][gr][c][nz] < 0)
    if kind == "burst":
        assert seen == {0, 1, 2, 3}, "transient input must exercise START/SHORT/STOP"


def test_chunking_does_not_change_bytes(oracle):
    l, r = make_signal("noise", 9 * 1152 + 500, 44100, 3)
    a, _, _ = oracle.encode_stream(2, 44100, 128, l, r)
    b, sz, _ = oracle.encode_stream(2, 44100, 128, l, r, chunk=777)
    c, _, _ = oracle.encode_stream(2, 44100, 128, l, r, chunk=4000)
    assert a == b == c
    assert sz[0] == 0 and sz[1] == 417        # 528 + 777 < 1904 <= 528 + 2*777: the second call completes frame 0


def test_flush_twice_and_reuse(oracle):
    enc = oracle.OracleEncoder(1, 44100, 128)
    x = make_signal("sine", 3000, 44100)[0]
    a = enc.encode_buffer(x) + enc.flush()
    assert enc.flush() == b""                 # Lame.js:1397-1399
    assert len(a) > 0
    b = enc.encode_buffer(x) + enc.flush()    # encoder stays usable
    assert len(b) > 0


def test_sine_peak_lands_in_expected_mdct_bin(oracle):
    n = 20 * 1152
    t = np.arange(n)
    s = np.rint(10000 * np.sin(2 * np.pi * 1000.0 * t / 44100)).astype(np.int16)
    _, _, tr = oracle.encode_stream(1, 44100, 128, s, None, trace_frames=30)
    x = tr["xr"][8, 0, 0]
    assert abs(int(np.argmax(np.abs(x))) - 1000 / (22050 / 576)) < 1.5   # MDCT phase moves the peak between bins 25 and 26
