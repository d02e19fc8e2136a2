"""SURVEY.md 8(f2), oracle side: joint stereo (four psycho-acoustic channels, M/S decision, ms_convert, reduce_side) and the bit
reservoir (header ring, main_data_begin back-pointers, build-up rule, drains) of oracle/ against bytes REAL lamejs produced in
those modes (tests/golden/lamejs_modes_golden.json, made by tests/golden/make_lamejs_modes_golden.py; `Mp3Encoder` itself fixes
STEREO and disable_reservoir, index.js:104,108 -- the fixtures come from the same module wiring with those two assignments
changed in the driver).  The CUDA library does not offer these modes."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_lamejs_modes_golden import stereo_signal  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "lamejs_modes_golden.json")))


def _encode(oracle, c, l, r):
    enc = oracle.OracleEncoder(c["channels"], c["samplerate"], c["kbps"], reservoir=not c["disable_reservoir"], joint_stereo=c["mode"] == "JOINT_STEREO")
    out, sizes = bytearray(), []
    step = c["chunk"] or max(len(l), 1)
    for i in range(0, len(l), step):
        b = enc.encode_buffer(l[i:i + step], None if r is None else r[i:i + step])
        sizes.append(len(b))
        out += b
    b = enc.flush()
    sizes.append(len(b))
    out += b
    crc, nb = enc.music_crc(), enc.bytes_written()
    enc.close()
    return bytes(out), sizes, crc, nb


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_lamejs_in_this_mode(oracle, name):
    c = GOLD[name]
    l, r = stereo_signal(c)
    data, sizes, crc, nb = _encode(oracle, c, l, r)
    assert len(data) == c["bytes"] and data[:16].hex() == c["head"]
    assert hashlib.sha256(data).hexdigest() == c["sha256"]
    assert len(sizes) == c["calls"] and hashlib.sha256(json.dumps([int(s) for s in sizes]).encode()).hexdigest() == c["sizes_sha256"]
    assert crc == c["music_crc"] and nb == c["bytes_written"]


def _frames(data):
    """(mode, mode_ext, main_data_begin) of every frame of a CBR MPEG-1 stream whose frames are contiguous"""
    out, i = [], 0
    while i + 6 <= len(data):
        assert data[i] == 0xFF and (data[i + 1] & 0xFE) == 0xFA, i
        kbps = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320][data[i + 2] >> 4]
        sr = [44100, 48000, 32000][(data[i + 2] >> 2) & 3]
        out.append(((data[i + 3] >> 6) & 3, (data[i + 3] >> 4) & 3, (data[i + 4] << 1) | (data[i + 5] >> 7)))
        i += 144000 * kbps // sr + ((data[i + 2] >> 1) & 1)
    return out


def test_the_fixtures_exercise_the_modes(oracle):
    """M/S frames and L/R frames both occur, the back-pointer is used, and the frame grid stays fixed with the reservoir on"""
    c = GOLD["joint_stereo_resv_sweep_2_44100_192_5000_swap"]
    data = _encode(oracle, c, *stereo_signal(c))[0]
    fr = _frames(data)
    assert all(m == 1 for m, _, _ in fr)                          # header says joint stereo
    exts = [x for _, x, _ in fr]
    assert exts.count(2) > 10 and exts.count(0) > 10              # mid/side frames and left/right frames
    mdb = [b for _, _, b in fr]
    assert mdb[0] == 0 and max(mdb) > 50 and max(mdb) <= 511      # main_data_begin back-pointers within their 9 bits
    c = GOLD["stereo_resv_noise_2_44100_128_1152_corr"]
    fr = _frames(_encode(oracle, c, *stereo_signal(c))[0])
    assert all(m == 0 and x == 0 for m, x, _ in fr) and max(b for _, _, b in fr) > 0
