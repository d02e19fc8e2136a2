#!/usr/bin/env python3
"""Randomised soak: oracle/ against REAL lamejs (tools/jsrun) on fresh random inputs.
usage: soak.py [ncases=64] [max_frames=300] [seed=1] [workers=6]   -> prints mismatches, exit code 1 if any."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(args):
    idx, seed, max_frames, native_only = args
    import oracle_lib as O
    import ref_lamejs as R
    from synth import make_signal
    rng = np.random.default_rng(seed * 100003 + idx)
    while True:
        ch = int(rng.integers(1, 3))
        sr = int(rng.choice([32000, 44100, 48000] if native_only else [8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000]))
        kbps = int(rng.choice([8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]))
        try:
            O.OracleEncoder(ch, sr, kbps).close()
            break
        except ValueError:
            continue
    kind = str(rng.choice(["noise", "white", "octave", "burst", "sweep", "sine"]))
    n = int(rng.integers(3, max_frames)) * 1152 + int(rng.integers(0, 1152))
    chunk = [None, 1152, int(rng.integers(1, 6000))][int(rng.integers(0, 3))]
    l, r = make_signal(kind, n, sr, 5000 + idx + seed * 7919)
    if kind in ("sweep", "sine"):   # vary level so that the ATH / analog-silence paths see quiet input too
        sh = int(rng.integers(0, 9))
        l, r = (l >> sh).astype(np.int16), (r >> sh).astype(np.int16)
    ref, ref_sizes, info = R.encode(ch, sr, kbps, l, r, chunk=chunk)
    got, sizes, _ = O.encode_stream(ch, sr, kbps, l, r if ch == 2 else None, chunk=chunk)
    ok = got == ref and sizes == ref_sizes
    first = -1
    if not ok:
        m = min(len(got), len(ref))
        first = next((i for i in range(m) if got[i] != ref[i]), m)
    return idx, ok, (ch, sr, kbps, kind, n, chunk), first, n // 1152


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    max_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    workers = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    import oracle_lib as O
    import ref_lamejs as R
    O.build(); R.build()
    bad, frames = 0, 0
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for idx, ok, cfg, first, nf in ex.map(one, [(i, seed, max_frames, True) for i in range(ncases)]):
            frames += nf
            if not ok:
                bad += 1
                print("MISMATCH case", idx, cfg, "first differing byte", first, flush=True)
    print("soak: %d cases, %d frames, %d mismatches (seed %d)" % (ncases, frames, bad, seed))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
