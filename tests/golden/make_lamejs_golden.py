#!/usr/bin/env python3
"""Generates tests/golden/lamejs_golden.json: SHA-256 / length / per-call sizes of the output of REAL lamejs --
the unmodified /root/reference sources executed by a real JavaScript engine (Qt QJSEngine, tools/jsrun/) -- for
seeded inputs of tests/synth.py.  These are the fixtures that pin oracle/ (and through it the CUDA path) to the
reference itself; they travel to the GPU box, the engine and /root/reference do not have to.

  python tests/golden/make_lamejs_golden.py            # all cases, 8 processes, a few minutes
Only cases lamejs itself can execute are recorded.  `cases()` is the single source of truth for the case list."""
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools", "jsrun"))
from synth import make_signal  # noqa: E402

BITRATES_V1 = [32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]


def cases():
    """name -> dict(kind, channels, samplerate, kbps, samples, seed, chunk)."""
    c = {}

    def add(name, kind, ch, sr, kbps, n, seed, chunk=0):
        c[name] = dict(kind=kind, channels=ch, samplerate=sr, kbps=kbps, samples=n, seed=seed, chunk=chunk)

    # the nine cases of golden.json (whole stream in one encodeBuffer call)
    add("c1_silence_mono_128", "silence", 1, 44100, 128, 44100, 0)
    add("c2_sweep_stereo_128", "sweep", 2, 44100, 128, 200 * 1152, 0)
    add("c3_white_stereo_48k_320", "white", 2, 48000, 320, 120 * 1152, 3)
    add("c4_octave_mono_128", "octave", 1, 44100, 128, 150 * 1152, 4)
    add("c5_burst_stereo_128", "burst", 2, 44100, 128, 150 * 1152, 5)
    add("noise_stereo_32k_192", "noise", 2, 32000, 192, 60 * 1152 + 17, 6)
    add("sine_mono_48k_256", "sine", 1, 48000, 256, 50 * 1152 + 901, 7)
    add("white_mono_44k_320", "white", 1, 44100, 320, 40 * 1152, 8)
    add("tiny_stereo_128", "noise", 2, 44100, 128, 10, 9)
    # the lamejs call pattern: README 1152-sample calls and odd chunkings (per-call byte counts are recorded)
    add("c1_silence_mono_128_chunk1152", "silence", 1, 44100, 128, 44100, 0, 1152)
    add("burst_stereo_128_chunk1152", "burst", 2, 44100, 128, 60 * 1152 + 5, 11, 1152)
    add("noise_stereo_128_chunk777", "noise", 2, 44100, 128, 40 * 1152 + 3, 12, 777)
    add("octave_mono_160_chunk5000", "octave", 1, 44100, 160, 50 * 1152, 13, 5000)
    # longer streams: more of the rate loop's rare branches
    add("burst_stereo_128_long", "burst", 2, 44100, 128, 600 * 1152, 21)
    add("white_stereo_48k_320_long", "white", 2, 48000, 320, 400 * 1152, 22)
    add("octave_mono_128_long", "octave", 1, 44100, 128, 600 * 1152, 23)
    add("sweep_stereo_128_long", "sweep", 2, 44100, 128, 1000 * 1152, 0)
    add("noise_stereo_192_long", "noise", 2, 44100, 192, 400 * 1152, 24)
    # configuration matrix: every sample rate x bitrate x channel count Mp3Encoder accepts, 24 frames each,
    # fed README style (1152-sample calls) so that the resampling configurations stay well defined in lamejs
    for sr in (8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000):
        for kbps in ([8, 16, 24] if sr < 32000 else []) + BITRATES_V1:
            for ch in (1, 2):
                kind = ("noise", "octave", "burst")[(kbps + ch) % 3]
                add("matrix_%d_%d_%d" % (sr, kbps, ch), kind, ch, sr, kbps, 24 * 1152 + 100, 100 + kbps + ch, 1152)
    return c


def run_case(item):
    name, c = item
    import ref_lamejs as R
    l, r = make_signal(c["kind"], c["samples"], c["samplerate"], c["seed"])
    try:
        data, sizes, info = R.encode(c["channels"], c["samplerate"], c["kbps"], l, r, chunk=c["chunk"] or None)
    except Exception as e:  # lamejs itself threw
        return name, dict(c, error=str(e)[-300:])
    return name, dict(c, bytes=len(data), sha256=hashlib.sha256(data).hexdigest(), head=data[:48].hex(),
                      sizes_sha256=hashlib.sha256(json.dumps(sizes).encode()).hexdigest(), calls=len(sizes),
                      nonempty_calls=sum(1 for s in sizes if s))


TAP_CASES = {   # name -> (kind, channels, samplerate, kbps, samples, seed): lamejs's own INTERMEDIATES are recorded for these
    "taps_burst_stereo_128": ("burst", 2, 44100, 128, 60 * 1152 + 5, 31),
    "taps_white_stereo_48k_320": ("white", 2, 48000, 320, 24 * 1152, 32),
    "taps_octave_mono_128": ("octave", 1, 44100, 128, 40 * 1152, 33),
    "taps_burst_mono_22k_32": ("burst", 1, 22050, 32, 80 * 576, 34),
    "taps_noise_stereo_16k_64": ("noise", 2, 16000, 64, 60 * 576 + 9, 35),
}
TAP_KEYS = ["xr", "en_l", "thm_l", "en_s", "thm_s", "blocktype", "ath_adjust", "l3_enc", "global_gain", "part2_3_length", "part2_length",
            "big_values", "count1", "scalefac_compress"]


def tap_hash(a):
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_tap_case(item):
    name, (kind, ch, sr, kbps, n, seed) = item
    import ref_lamejs as R
    l, r = make_signal(kind, n, sr, seed)
    data, taps = R.encode_with_taps(ch, sr, kbps, l, r)
    return name, dict(kind=kind, channels=ch, samplerate=sr, kbps=kbps, samples=n, seed=seed, bytes=len(data),
                      sha256=hashlib.sha256(data).hexdigest(), frames=int(taps["xr"].shape[0]), granules=int(taps["xr"].shape[1]),
                      taps={k: tap_hash(taps[k]) for k in TAP_KEYS})


def main():
    import ref_lamejs as R
    assert R.available(), "needs /root/reference and the Qt JS engine of this image"
    R.build()
    cs = cases()
    out = {}
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for name, res in ex.map(run_case, cs.items()):
            out[name] = res
            print(name, res.get("bytes"), res.get("error", ""), flush=True)
    taps = {}
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for name, res in ex.map(run_tap_case, TAP_CASES.items()):
            taps[name] = res
            print(name, res["frames"], "frames of intermediates", flush=True)
    meta = {"_generator": "tests/golden/make_lamejs_golden.py", "_engine": "Qt QJSEngine 6.6.3 (libQt6Qml), Math.* = C libm",
            "_reference_commit": json.load(open("/root/reference/.SUBMODULES.json"))["commit"], "_loader": "lame.all.js"}
    json.dump({"meta": meta, "cases": out, "taps": taps}, open(os.path.join(HERE, "lamejs_golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "fixtures")


if __name__ == "__main__":
    main()
