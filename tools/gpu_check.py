#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle diagnostic (run on the GPU box; prints where parity first breaks)."""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import lamejs_b200 as M  # noqa: E402
from synth import make_signal  # noqa: E402


def first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.nonzero(a.view(np.uint32 if a.dtype == np.float32 else a.dtype) != b.view(np.uint32 if b.dtype == np.float32 else b.dtype))
    if len(d[0]) == 0:
        return None
    idx = tuple(int(x[0]) for x in d)
    return idx, a[idx], b[idx], len(d[0])


def check(name, ch, sr, kbps, l, r, verbose=True):
    F = M.stream_frames(len(l))
    data, sizes, tr = O.encode_stream(ch, sr, kbps, l, r, trace_frames=F + 2)
    assert len(tr) == F, (len(tr), F)
    ok = True
    # stage 1: MDCT with the oracle's block types
    fb = tr["blocktype"][:, :, :ch].astype(np.int32)
    g = M.debug_stages(ch, sr, kbps, l, r, force_blocktype=fb, want=("xr",))
    d = first_diff(g["xr"], tr["xr"][:, :, :ch])
    print("[%s] xr (forced blocktype): %s" % (name, "OK" if d is None else "DIFF %s" % (d,)))
    ok &= d is None
    t0 = time.time()
    g = M.debug_stages(ch, sr, kbps, l, r, want=("xr", "blocktype", "en_l", "thm_l", "en_s", "thm_s", "ath_adjust", "l3_enc", "ginfo", "bytes"))
    t1 = time.time()
    for key, ref in [("blocktype", tr["blocktype"][:, :, :ch]), ("ath_adjust", tr["ath_adjust"]),
                     ("en_l", tr["en_l"][:, :, :ch]), ("thm_l", tr["thm_l"][:, :, :ch]),
                     ("en_s", tr["en_s"][:, :, :ch]), ("thm_s", tr["thm_s"][:, :, :ch]), ("xr", tr["xr"][:, :, :ch])]:
        a = g[key]
        if key == "ath_adjust":
            dd = np.nonzero(a != ref)[0]
            d = None if len(dd) == 0 else (int(dd[0]), a[dd[0]], ref[dd[0]], len(dd))
        else:
            d = first_diff(a, ref.astype(a.dtype))
        print("[%s] %s: %s" % (name, key, "OK" if d is None else "DIFF %s" % (d,)))
        ok &= d is None
    gi = g["ginfo"]
    for j, key in enumerate(["global_gain", "part2_3_length", "part2_length", "big_values", "count1", "scalefac_compress"]):
        ref = tr[key][:, :, :ch]
        dd = np.argwhere(gi[..., j] != ref)
        print("[%s] %s: %s" % (name, key, "OK" if len(dd) == 0 else "DIFF first %s gpu %s ref %s (n=%d)" % (dd[0], gi[..., j][tuple(dd[0])], ref[tuple(dd[0])], len(dd))))
        ok &= len(dd) == 0
    dd = np.argwhere(g["l3_enc"] != tr["l3_enc"][:, :, :ch])
    print("[%s] l3_enc: %s" % (name, "OK" if len(dd) == 0 else "DIFF first %s (n=%d)" % (dd[0], len(dd))))
    ok &= len(dd) == 0
    gb = g["bytes"].tobytes()
    if gb == data:
        print("[%s] bytes: OK (%d bytes, %d frames) gpu %.3fs" % (name, len(data), F, t1 - t0))
    else:
        n = min(len(gb), len(data))
        k = next((i for i in range(n) if gb[i] != data[i]), n)
        print("[%s] bytes: DIFF at byte %d of %d (len gpu %d)" % (name, k, len(data), len(gb)))
        ok = False
    return ok


def main():
    cases = [
        ("noise-st-128", 2, 44100, 128, "noise", 30),
        ("sine-mono-128", 1, 44100, 128, "sine", 20),
        ("burst-st-128", 2, 44100, 128, "burst", 40),
        ("white-st-320-48k", 2, 48000, 320, "white", 30),
        ("silence-mono", 1, 44100, 128, "silence", 10),
        ("sweep-st-128", 2, 44100, 128, "sweep", 60),
    ]
    allok = True
    for name, ch, sr, kbps, kind, frames in cases:
        l, r = make_signal(kind, frames * 1152 + 77, sr, seed=zlib.crc32(name.encode()) & 0xffff)
        allok &= check(name, ch, sr, kbps, l, r if ch == 2 else None)
    print("ALL OK" if allok else "SOME DIFFS")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
