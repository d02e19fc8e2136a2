#!/usr/bin/env python3
"""Minimal driver for ncu: encodes the C2 workload (or a prefix) N times from device-resident PCM."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import lamejs_b200 as M  # noqa: E402
from synth import make_signal  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = frames * 1152
l, r = make_signal("sweep", 10000 * 1152, 44100)
l, r = l[:n], r[:n]
pcm = torch.from_numpy(np.concatenate([l, r])).cuda()
nb = M.stream_bytes(2, 44100, 128, n)
out = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    tm = M.encode_streams_device(2, 44100, 128, pcm.data_ptr(), [0], [n], out.data_ptr(), [0])
torch.cuda.synchronize()
print("timings ms [psy, scan, mask, fb, q1, qn, total, passes | prepare, search, outer, finish, pack]:", [round(float(x), 3) for x in tm[:13]])
import hashlib
print("sha256:", hashlib.sha256(out[:nb].cpu().numpy().tobytes()).hexdigest()[:16], "lib:", os.environ.get("MP3B200_LIB", "default"))
print("x realtime:", (M.stream_frames(n) * 1152 / 44100) / (tm[6] / 1000))

if hasattr(M.lib(), "mp3b200_debug_qstats"):
    import ctypes
    st = (ctypes.c_ulonglong * 16)()
    M.lib().mp3b200_debug_qstats(st, 1)
    names = ["bs1 search", "bs1 tail", "reval-gr0 search", "reval-gr0 tail", "reval-gr1 search", "reval packs", "divide_init r0 calls", "divide_init r1 calls", "outer huff loop",
             "outer best loop", "calc_noise(outer)", "gc encoded", "reval gr0", "reval gr1", "reval rate loops (gc)", "noquant region calls"]
    gcs = max(1, st[11])
    print("call counters over %d encodes (per encoded gc in brackets):" % reps)
    for n_, v in zip(names, st):
        if n_ != "-": print("  %-28s %10d  [%.2f]" % (n_, v, v / gcs))
