#!/usr/bin/env python3
"""Locate the first frame where GPU and oracle differ on a long stream, and whether the GPU is self-consistent."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import lamejs_b200 as M
from synth import make_signal

kind, frames = (sys.argv[1] if len(sys.argv) > 1 else "sweep"), int(sys.argv[2]) if len(sys.argv) > 2 else 10000
ch, sr, kbps = 2, 44100, 128
l, r = make_signal(kind, frames * 1152, sr)
F = M.stream_frames(len(l))
ref, _, tr = O.encode_stream(ch, sr, kbps, l, r, trace_frames=F + 2)
a = M.encode_streams(ch, sr, kbps, [l], [r])[0]
b = M.encode_streams(ch, sr, kbps, [l], [r])[0]
print("gpu run1 == run2:", a == b, " gpu == oracle:", a == ref)
g = M.debug_stages(ch, sr, kbps, l, r, want=("xr", "blocktype", "en_l", "thm_l", "en_s", "thm_s", "ath_adjust", "l3_enc", "ginfo", "bytes"))
print("debug_stages bytes == oracle:", g["bytes"].tobytes() == ref, " == run1:", g["bytes"].tobytes() == a)
def fd(name, x, y):
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
    if x.dtype == np.float32: x = x.view(np.uint32); y = y.view(np.uint32)
    d = np.argwhere(x != y)
    print("%-12s %s" % (name, "OK" if len(d) == 0 else "first diff at %s (n=%d)" % (d[0], len(d))))
fd("blocktype", g["blocktype"], tr["blocktype"][:, :, :ch]); fd("ath", g["ath_adjust"], tr["ath_adjust"])
for k in ("xr", "en_l", "thm_l", "en_s", "thm_s"): fd(k, g[k], tr[k][:, :, :ch])
fd("l3_enc", g["l3_enc"], tr["l3_enc"][:, :, :ch])
for j, k in enumerate(["global_gain", "part2_3_length", "part2_length", "big_values", "count1", "scalefac_compress"]): fd(k, g["ginfo"][..., j], tr[k][:, :, :ch])
# frame-level byte diff
off = 0; nd = 0
for k in range(F):
    fl = int(tr["frame_bytes"][k])
    if a[off:off+fl] != ref[off:off+fl]:
        if nd < 5: print("frame", k, "differs; oracle old_value in/out", tr["old_value_in"][k], tr["old_value_out"][k], "steps", tr["cur_step_in"][k], tr["cur_step_out"][k], "gg gpu", g["ginfo"][k, :, :, 0].tolist(), "ref", tr["global_gain"][k].tolist())
        nd += 1
    off += fl
print("differing frames:", nd)
import mp3_parse
books = mp3_parse.HuffBooks(O.TABLES_H)
off = 0
for k in range(F):
    fl = int(tr["frame_bytes"][k])
    if a[off:off+fl] != ref[off:off+fl]:
        fo = mp3_parse.parse_frame(ref, off, books)
        try:
            fg = mp3_parse.parse_frame(a, off, books)
        except Exception as e:
            print("GPU frame", k, "does not parse:", repr(e)); fg = None
        if fg:
            print("scfsi", fo["scfsi"], fg["scfsi"])
            for gr in range(2):
                for c in range(ch):
                    x, y = fo["gi"][gr][c], fg["gi"][gr][c]
                    for key in x:
                        if key == "ix":
                            if not np.array_equal(x[key], y[key]): print(" gr", gr, "ch", c, "ix differs at", np.nonzero(x[key] != y[key])[0][:10])
                        elif x[key] != y[key]:
                            print(" gr", gr, "ch", c, key, "oracle", x[key], "gpu", y[key])
        nb = [i for i in range(fl) if a[off+i] != ref[off+i]]
        print("frame", k, "byte diffs at", nb[:20], "of", fl)
        break
    off += fl
