#!/usr/bin/env python3
"""Re-makes committed lamejs fixtures with tools/jsrun/fdlibm.js loaded first (the engine's Math.log / log10 / exp / pow then are
fdlibm's, as under V8) and compares with what the engine's C library produced: if the SHA-256s agree, the bytes do not depend on
which of the two libms computed those functions.
usage: fdlibm_check.py [every=6] [workers=8]   -> every n-th of the 306 Mp3Encoder fixtures; exit code 1 on a difference"""
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def one(item):
    import ref_lamejs as R
    from synth import make_signal
    name, c = item
    l, r = make_signal(c["kind"], c["samples"], c["samplerate"], seed=c["seed"])
    data, sizes, _ = R.encode(c["channels"], c["samplerate"], c["kbps"], l, r if c["channels"] == 2 else None, chunk=c["chunk"] or None, fdlibm=True)
    return name, hashlib.sha256(data).hexdigest() == c["sha256"] and len(data) == c["bytes"]


def main():
    every = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    from make_lamejs_golden import cases
    fix = json.load(open(os.path.join(ROOT, "tests", "golden", "lamejs_golden.json")))["cases"]
    spec = cases()
    names = [n for n in sorted(fix) if "error" not in fix[n] and fix[n]["bytes"] < 600000][::every]
    items = [(n, dict(spec[n], sha256=fix[n]["sha256"], bytes=fix[n]["bytes"])) for n in names]
    bad = 0
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for name, ok in ex.map(one, items):
            if not ok:
                bad += 1
                print("DIFFERENT under fdlibm:", name, flush=True)
    print("fdlibm_check: %d fixtures re-made with fdlibm's log / log10 / exp / pow, %d differ" % (len(items), bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
