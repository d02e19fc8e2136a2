#!/usr/bin/env python3
"""Generates tests/golden/lamejs_modes_golden.json: bytes REAL lamejs produces in the modes `Mp3Encoder` hard-codes away --
gfp.mode = JOINT_STEREO and / or gfp.disable_reservoir = false; unmodified /root/reference under Qt's
QJSEngine with index.js's module wiring and those two assignments changed in the DRIVER (tools/jsrun/mode_probe.py).

  python tests/golden/make_lamejs_modes_golden.py      # ~1 minute, 8 processes"""
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools", "jsrun"))
from synth import make_signal  # noqa: E402


def stereo_signal(c):
    """`image`: indep = the two generator channels; corr = right is 3/4 left + 1/4 right (M/S pays); swap = first half
    correlated, second half independent (the M/S decision changes inside the stream)."""
    l, r = make_signal(c["kind"], c["samples"], c["samplerate"], seed=c["seed"])
    if c["channels"] == 1:
        return l, None
    mix = ((l.astype(np.int32) * 3 + r.astype(np.int32)) // 4).astype(np.int16)
    if c["image"] == "corr":
        r = mix
    elif c["image"] == "swap":
        h = len(l) // 2
        r = np.concatenate([mix[:h], r[h:]])
    return l, r


def cases():
    c = {}

    def add(kind, ch, sr, kbps, n, seed, chunk, image, mode, nores):
        name = "%s_%s_%s_%d_%d_%d_%d_%s" % (mode.lower(), "nores" if nores else "resv", kind, ch, sr, kbps, chunk, image)
        c[name] = dict(kind=kind, channels=ch, samplerate=sr, kbps=kbps, samples=n, seed=seed, chunk=chunk, image=image, mode=mode, disable_reservoir=nores)

    for mode, nores in (("STEREO", False), ("JOINT_STEREO", True), ("JOINT_STEREO", False)):
        add("noise", 2, 44100, 128, 40 * 1152 + 100, 61, 1152, "corr", mode, nores)
        add("burst", 2, 44100, 128, 60 * 1152 + 5, 62, 0, "swap", mode, nores)
        add("white", 2, 48000, 320, 30 * 1152, 63, 1152, "corr", mode, nores)
        add("octave", 2, 32000, 96, 30 * 1152 + 17, 64, 777, "indep", mode, nores)
        add("sweep", 2, 44100, 192, 80 * 1152, 65, 5000, "swap", mode, nores)
        add("noise", 2, 22050, 64, 50 * 576 + 9, 66, 576, "corr", mode, nores)      # MPEG-2
        add("burst", 2, 11025, 32, 50 * 576, 67, 1152, "swap", mode, nores)          # MPEG-2.5
        add("silence", 2, 44100, 128, 20 * 1152, 0, 1152, "indep", mode, nores)
    add("octave", 1, 44100, 96, 40 * 1152, 68, 1152, "indep", "STEREO", False)       # mono with the reservoir
    add("noise", 1, 16000, 32, 50 * 576, 69, 576, "indep", "STEREO", False)
    add("burst", 2, 44100, 128, 300 * 1152, 70, 4096, "swap", "JOINT_STEREO", False)  # a long one
    return c


def _run(item):
    import mode_probe as P
    name, c = item
    l, r = stereo_signal(c)
    data, sizes, o = P.encode(c["channels"], c["samplerate"], c["kbps"], l, r, chunk=c["chunk"] or None, mode=c["mode"],
                              disable_reservoir=c["disable_reservoir"])
    return name, dict(c, bytes=len(data), sha256=hashlib.sha256(data).hexdigest(), calls=len(sizes),
                      sizes_sha256=hashlib.sha256(json.dumps([int(s) for s in sizes]).encode()).hexdigest(),
                      music_crc=o["crc"], bytes_written=o["nbytes"], head=data[:16].hex())


def main():
    out = {}
    with ProcessPoolExecutor(8) as ex:
        for name, r in ex.map(_run, cases().items()):
            out[name] = r
    with open(os.path.join(HERE, "lamejs_modes_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
