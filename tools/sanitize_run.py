#!/usr/bin/env python3
"""Small workloads for compute-sanitizer (racecheck / memcheck / initcheck): transient bursts (block switching), 320 kbps
white noise (long rate loops), an MPEG-2 stream, live handles, state export / import / seek.  Every result is checked against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lamejs_b200 as M  # noqa: E402
import oracle_lib as O  # noqa: E402
from synth import make_signal  # noqa: E402

cases = [(2, 44100, 128, "burst", 24 * 1152 + 5), (2, 48000, 320, "white", 16 * 1152), (1, 44100, 128, "octave", 20 * 1152),
         (2, 22050, 64, "burst", 30 * 576 + 9), (1, 8000, 16, "noise", 24 * 576)]
bad = 0
for ch, sr, kbps, kind, n in cases:
    sig = [make_signal(kind, n - 300 * j, sr, 50 + j) for j in range(3)]
    outs = M.encode_streams(ch, sr, kbps, [s[0] for s in sig], [s[1] for s in sig] if ch == 2 else None)
    for s, o in zip(sig, outs):
        ok = o == O.encode_stream(ch, sr, kbps, s[0], s[1] if ch == 2 else None)[0]
        bad += 0 if ok else 1
    l, r = sig[0]
    enc = M.Mp3Encoder(ch, sr, kbps)
    got = b"".join(enc.encodeBuffer(l[i:i + 1152], r[i:i + 1152] if ch == 2 else None) for i in range(0, n, 1152)) + enc.flush()
    enc.close()
    ok = got == O.encode_stream(ch, sr, kbps, l, r if ch == 2 else None, chunk=1152)[0]
    bad += 0 if ok else 1
    # encoder state: checkpoint / resume in the middle, and the stream cut into three frame ranges with a warm-up
    from lamejs_b200 import sharding
    fs = 576 * M.granules_per_frame(ch, sr, kbps)
    want = O.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)[0]
    a = M.Mp3Encoder(ch, sr, kbps)
    got = a.encodeBuffer(l[:n // 2], r[:n // 2] if ch == 2 else None)
    blob = a.export_state(); a.close()
    b = M.Mp3Encoder(ch, sr, kbps); b.import_state(blob)
    got += b.encodeBuffer(l[n // 2:], r[n // 2:] if ch == 2 else None) + b.flush(); b.close()
    bad += 0 if got == want else 1
    got, _ = sharding.encode_stream_segments_local(lambda: M.Mp3Encoder(ch, sr, kbps), l, r if ch == 2 else None, fs, 3, 4)
    bad += 0 if got == want else 1
print("sanitize_run: %d cases, %d mismatches" % (len(cases) * 6, bad))
sys.exit(1 if bad else 0)
