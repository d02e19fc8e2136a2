#!/usr/bin/env python3
"""Generates tests/golden/golden.json: SHA-256 + length + first bytes of the ORACLE's MP3 output for small seeded
inputs.  lamejs has no golden vectors and cannot run in this image (no JS engine), so these fixtures pin the oracle
(against regressions) and give the GPU tests a reference that does not need the oracle at run time.  Re-run only when
the oracle changes on purpose:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from synth import make_signal  # noqa: E402

CASES = [  # name, kind, channels, samplerate, kbps, samples, seed
    ("c1_silence_mono_128", "silence", 1, 44100, 128, 44100, 0),
    ("c2_sweep_stereo_128", "sweep", 2, 44100, 128, 200 * 1152, 0),
    ("c3_white_stereo_48k_320", "white", 2, 48000, 320, 120 * 1152, 3),
    ("c4_octave_mono_128", "octave", 1, 44100, 128, 150 * 1152, 4),
    ("c5_burst_stereo_128", "burst", 2, 44100, 128, 150 * 1152, 5),
    ("noise_stereo_32k_192", "noise", 2, 32000, 192, 60 * 1152 + 17, 6),
    ("sine_mono_48k_256", "sine", 1, 48000, 256, 50 * 1152 + 901, 7),
    ("white_mono_44k_320", "white", 1, 44100, 320, 40 * 1152, 8),
    ("tiny_stereo_128", "noise", 2, 44100, 128, 10, 9),
]


def main():
    out = {}
    for name, kind, ch, sr, kbps, n, seed in CASES:
        l, r = make_signal(kind, n, sr, seed)
        data, _, _ = O.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)
        out[name] = {"kind": kind, "channels": ch, "samplerate": sr, "kbps": kbps, "samples": n, "seed": seed,
                     "bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(), "head": data[:48].hex()}
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "fixtures")


if __name__ == "__main__":
    main()
