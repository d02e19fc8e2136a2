#!/usr/bin/env python3
"""Randomised parity soak: many (config, signal, length) cases, GPU batch API against the CPU oracle, byte for byte.
usage: gpu_soak.py [cases] [seed]"""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lamejs_b200 as M
import oracle_lib as O
from synth import make_signal

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 160
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20260923)
O.lib()
kinds = ["noise", "burst", "sweep", "white", "sine", "octave", "silence"]
configs = [(ch, sr, kbps) for sr in (32000, 44100, 48000) for kbps in (64, 96, 112, 128, 160, 192, 256, 320) for ch in (1, 2)
           if M.stream_bytes(ch, sr, kbps, 1152) > 0]
cases = []
for i in range(ncases):
    ch, sr, kbps = configs[rng.integers(len(configs))]
    kind = kinds[rng.integers(len(kinds))]
    n = int(rng.integers(1, 420 * 1152))
    cases.append((ch, sr, kbps, kind, n, int(rng.integers(1 << 30))))
t0 = time.time()
sigs = [make_signal(k, n, sr, seed) for (ch, sr, kbps, k, n, seed) in cases]
def ref(i):
    ch, sr, kbps = cases[i][:3]
    l, r = sigs[i]
    return O.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)[0]
with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
    fut = [ex.submit(ref, i) for i in range(ncases)]
    bad = 0
    by_cfg = {}
    for i, c in enumerate(cases): by_cfg.setdefault(c[:3], []).append(i)
    got = [None] * ncases
    for (ch, sr, kbps), idx in by_cfg.items():
        outs = M.encode_streams(ch, sr, kbps, [sigs[i][0] for i in idx], [sigs[i][1] for i in idx] if ch == 2 else None)
        for i, o in zip(idx, outs): got[i] = o
    for i in range(ncases):
        if got[i] != fut[i].result():
            bad += 1
            print("MISMATCH case", i, cases[i], flush=True)
frames = sum(M.stream_frames(c[4]) for c in cases)
print("soak: %d cases, %d configs, %d frames, %d mismatches, %.1f s" % (ncases, len(by_cfg), frames, bad, time.time() - t0))
sys.exit(1 if bad else 0)
