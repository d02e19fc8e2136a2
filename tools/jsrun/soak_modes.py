#!/usr/bin/env python3
"""Randomised soak for the modes and the tag row: oracle/ against REAL lamejs run in
STEREO / JOINT_STEREO x reservoir on / off (mode_probe.py), comparing bytes, per-call sizes, gfc.nMusicCRC and nBytesWritten.
usage: soak_modes.py [ncases=48] [max_frames=200] [seed=1] [workers=8]   -> prints mismatches, exit code 1 if any."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(args):
    idx, seed, max_frames = args
    import mode_probe as P
    import oracle_lib as O
    from synth import make_signal
    rng = np.random.default_rng(seed * 100003 + idx)
    while True:
        sr = int(rng.choice([8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000]))
        kbps = int(rng.choice([8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]))
        if O.out_samplerate(2, sr, kbps) == sr:
            break
    joint, nores = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    if not joint and nores:
        nores = False                       # plain Mp3Encoder is covered by soak.py
    kind = str(rng.choice(["noise", "white", "octave", "burst", "sweep", "sine"]))
    fs = 1152 if sr >= 32000 else 576
    n = int(rng.integers(3, max_frames)) * fs + int(rng.integers(0, fs))
    chunk = [None, 1152, int(rng.integers(1, 6000))][int(rng.integers(0, 3))]
    l, r = make_signal(kind, n, sr, 9000 + idx + seed * 7919)
    w = int(rng.integers(0, 5))             # stereo image: 0 = independent ... 4 = identical channels
    r = ((l.astype(np.int32) * w + r.astype(np.int32) * (4 - w)) // 4).astype(np.int16)
    if kind in ("sweep", "sine"):
        sh = int(rng.integers(0, 9))
        l, r = (l >> sh).astype(np.int16), (r >> sh).astype(np.int16)
    ref, ref_sizes, o = P.encode(2, sr, kbps, l, r, chunk=chunk, mode="JOINT_STEREO" if joint else "STEREO", disable_reservoir=nores)
    enc = O.OracleEncoder(2, sr, kbps, reservoir=not nores, joint_stereo=joint)
    out, sizes = bytearray(), []
    step = chunk or max(n, 1)
    for i in range(0, n, step):
        b = enc.encode_buffer(l[i:i + step], r[i:i + step])
        sizes.append(len(b))
        out += b
    b = enc.flush()
    sizes.append(len(b))
    out += b
    ok = bytes(out) == ref and sizes == ref_sizes and enc.music_crc() == o["crc"] and enc.bytes_written() == o["nbytes"]
    first = -1
    if not ok:
        m = min(len(out), len(ref))
        first = next((i for i in range(m) if out[i] != ref[i]), m)
    return idx, ok, (sr, kbps, "joint" if joint else "stereo", "nores" if nores else "resv", kind, n, chunk, w), first, n // fs


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    max_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    workers = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    import oracle_lib as O
    import ref_lamejs as R
    O.build(); R.build()
    bad, frames = 0, 0
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for idx, ok, cfg, first, nf in ex.map(one, [(i, seed, max_frames) for i in range(ncases)]):
            frames += nf
            if not ok:
                bad += 1
                print("MISMATCH case", idx, cfg, "first differing byte", first, flush=True)
    print("soak_modes: %d cases, %d frames, %d mismatches (seed %d)" % (ncases, frames, bad, seed))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
