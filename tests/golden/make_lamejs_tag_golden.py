#!/usr/bin/env python3
"""Generates tests/golden/lamejs_tag_golden.json: what REAL lamejs (unmodified /root/reference under Qt's QJSEngine,
tools/jsrun/tag_probe.py) computes for the container / metadata row, SURVEY.md 8(f3):

  hot     gfc.nMusicCRC and VBR_seek_table.nBytesWritten after an ordinary `lamejs.Mp3Encoder` run -- copy_buffer keeps
          them on every call (BitStream.js:924-935), tag or no tag
  tagged  the stream, the seek-table state and the buffer getLameTagFrame fills when gfp.bWriteVbrTag is true (unbound
          names of VBRTag.js bound by the probe, source text untouched)
  wav     lamejs.WavHeader.readHeader on hand-made RIFF byte strings

  python tests/golden/make_lamejs_tag_golden.py      # ~2 minutes, 8 processes
The fixtures travel to the GPU box; the engine and /root/reference do not."""
import hashlib
import json
import os
import struct
import sys
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools", "jsrun"))
from synth import make_signal  # noqa: E402


def hot_cases():
    c = {}

    def add(kind, ch, sr, kbps, n, seed, chunk):
        c["hot_%s_%d_%d_%d_%d" % (kind, ch, sr, kbps, chunk)] = dict(kind=kind, channels=ch, samplerate=sr, kbps=kbps, samples=n, seed=seed, chunk=chunk)

    add("noise", 2, 44100, 128, 40 * 1152 + 100, 31, 1152)
    add("burst", 2, 44100, 128, 60 * 1152 + 5, 32, 0)
    add("octave", 1, 44100, 128, 50 * 1152, 33, 5000)
    add("white", 2, 48000, 320, 30 * 1152, 34, 1152)
    add("sweep", 2, 32000, 192, 30 * 1152 + 17, 35, 777)
    add("noise", 1, 22050, 32, 40 * 576 + 9, 36, 1152)       # MPEG-2
    add("octave", 2, 24000, 64, 40 * 576, 37, 576)
    add("burst", 1, 8000, 8, 40 * 576, 38, 1152)             # MPEG-2.5
    add("noise", 2, 11025, 32, 40 * 576 + 1, 39, 1000)
    add("silence", 1, 44100, 128, 44100, 0, 1152)
    add("noise", 2, 44100, 128, 10, 40, 0)                   # nothing but the flush
    return c


def tagged_cases():
    c = {}

    def add(kind, ch, sr, kbps, n, seed, chunk):
        c["tag_%s_%d_%d_%d_%d" % (kind, ch, sr, kbps, chunk)] = dict(kind=kind, channels=ch, samplerate=sr, kbps=kbps, samples=n, seed=seed, chunk=chunk)

    add("noise", 2, 48000, 128, 24 * 1152 + 100, 41, 1152 * 4)       # integer frame sizes: JavaScript == Java byte for byte
    add("burst", 2, 48000, 320, 150 * 1152, 42, 0)
    add("octave", 1, 32000, 64, 40 * 1152 + 3, 43, 1152)
    add("white", 2, 32000, 256, 398 * 1152 - 800, 44, 0)             # 399 frames: the last fill level before the bag halves
    add("white", 2, 32000, 160, 420 * 1152, 53, 0)                   # > 400 frames: Java halves the bag, `bag[i / 2]` in JavaScript does not
    add("noise", 1, 48000, 96, 30 * 1152, 45, 5000)
    add("noise", 2, 24000, 64, 60 * 576 + 7, 46, 1152)               # MPEG-2
    add("octave", 1, 16000, 32, 50 * 576, 47, 576)
    add("noise", 1, 8000, 24, 50 * 576, 48, 1152)                    # MPEG-2.5, frame 216 bytes
    add("noise", 1, 8000, 8, 30 * 576, 49, 1152)                     # 72-byte frames: InitVbrTag switches the tag off
    add("sweep", 2, 44100, 128, 50 * 1152, 50, 1152)                 # fractional TotalFrameSize in JavaScript
    add("noise", 1, 22050, 48, 40 * 576, 51, 1152)
    add("burst", 2, 44100, 320, 100 * 1152 + 11, 52, 0)
    return c


def wav_cases():
    def fmt(ch, sr, n=16, tag=1):
        body = struct.pack("<HHIIHH", tag, ch, sr, sr * ch * 2, ch * 2, 16)
        body += b"\0" * (n - 16)
        return b"fmt " + struct.pack("<I", n) + body

    def riff(chunks, form=b"WAVE", magic=b"RIFF"):
        body = form + b"".join(chunks)
        return magic + struct.pack("<I", len(body)) + body

    def data(n):
        return b"data" + struct.pack("<I", n) + bytes(range(256)) * (n // 256) + bytes(n % 256)

    lst = b"LIST" + struct.pack("<I", 26) + b"INFOISFT" + struct.pack("<I", 14) + b"Lavf58.29.100\0"
    c = {
        "pcm16_stereo_44k": riff([fmt(2, 44100), data(400)]),
        "pcm16_mono_8k_fmt18": riff([fmt(1, 8000, 18), data(100)]),
        "list_before_data": riff([fmt(2, 48000), lst, data(64)]),
        "two_chunks_before_data": riff([fmt(1, 22050), lst, b"fact" + struct.pack("<II", 4, 1234), data(10)]),
        "data_len_zero": riff([fmt(2, 32000), data(0)]),
        "data_len_larger_than_file": riff([fmt(2, 44100), b"data" + struct.pack("<I", 1 << 20) + b"\1\2\3\4"]),
        "not_riff": riff([fmt(2, 44100), data(16)], magic=b"RIFX"),
        "not_wave": riff([fmt(2, 44100), data(16)], form=b"AVI "),
        "fmt_not_first": riff([lst, fmt(2, 44100), data(16)]),
        "fmt_extensible_40": riff([fmt(2, 44100, 40, 0xFFFE), data(16)]),
        "no_data_chunk": riff([fmt(2, 44100), lst]),
        "truncated_header": riff([fmt(2, 44100), data(16)])[:22],
        "channels_6_96k": riff([fmt(6, 96000), data(24)]),
        "empty": b"",
    }
    return c


def _sig(c):
    return make_signal(c["kind"], c["samples"], c["samplerate"], seed=c["seed"])


def _run_hot(item):
    import tag_probe as T
    name, c = item
    l, r = _sig(c)
    data, crc, nb = T.hot_path_crc(c["channels"], c["samplerate"], c["kbps"], l, r if c["channels"] == 2 else None, chunk=c["chunk"] or None)
    return name, dict(c, bytes=len(data), sha256=hashlib.sha256(data).hexdigest(), music_crc=crc, bytes_written=nb)


def _run_tagged(item):
    import tag_probe as T
    name, c = item
    l, r = _sig(c)
    o = T.tagged(c["channels"], c["samplerate"], c["kbps"], l, r if c["channels"] == 2 else None, chunk=c["chunk"] or None)
    keep = {k: o[k] for k in ("rc", "sizes", "tag_ret", "frames", "pos", "sum", "want", "total_frame_size", "write_tag", "encoder_padding",
                              "lowpassfreq", "noise_shaping", "preset", "vbr_q", "exp_nspsytune", "athtype", "sideinfo_len", "mode_ext")}
    n = int(-(-o["total_frame_size"] // 1)) if o["write_tag"] else 0
    return name, dict(c, bytes=len(o["bytes"]), sha256=hashlib.sha256(o["bytes"]).hexdigest(), music_crc=o["crc"], bytes_written=o["nbytes"],
                      bag_sha256=hashlib.sha256(json.dumps(o["bag"]).encode()).hexdigest(), tag=o["tag"][:n].hex(),
                      first_bytes=o["bytes"][:8].hex(), **keep)


def _run_wav(item):
    import tag_probe as T
    name, b = item
    return name, dict(hex=b.hex(), result=T.wav_header(b))


def main():
    out = {"hot": {}, "tagged": {}, "wav": {}}
    with ProcessPoolExecutor(8) as ex:
        for name, r in ex.map(_run_hot, hot_cases().items()):
            out["hot"][name] = r
        for name, r in ex.map(_run_tagged, tagged_cases().items()):
            out["tagged"][name] = r
        for name, r in ex.map(_run_wav, wav_cases().items()):
            out["wav"][name] = r
    with open(os.path.join(HERE, "lamejs_tag_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
