"""Synthetic PCM generators: counter-based, integer-only where possible, so that every run and
every machine produces identical Int16 input for the oracle and the GPU encoder."""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def h64(seed, i):
    """splitmix64 finaliser of seed + (i+1)*golden (uint64 wrap)."""
    with np.errstate(over="ignore"):
        return _mix(np.uint64(seed) + (np.asarray(i, dtype=np.uint64) + np.uint64(1)) * _G)


def s16(seed, i):
    return (h64(seed, i) >> np.uint64(48)).astype(np.uint16).view(np.int16)


def white(n, seed, offset=0):
    i = np.arange(offset, offset + n, dtype=np.uint64)
    return s16(seed, i).copy(), s16(seed ^ 0xFFFF0000, i).copy()


def octave_hold(n, seed, offset=0):
    i = np.arange(offset, offset + n, dtype=np.uint64)
    x = (s16(seed, i) >> 1).astype(np.int32) + (s16(seed + 1, i >> np.uint64(1)) >> 2) + (s16(seed + 2, i >> np.uint64(2)) >> 3) + \
        (s16(seed + 3, i >> np.uint64(3)) >> 4)
    return x.astype(np.int16)


def bursts(n, seed, offset=0):
    i = np.arange(offset, offset + n, dtype=np.uint64)
    on = (i % np.uint64(4099)) < np.uint64(64)

    def chan(sd):
        dither = (((h64(sd, i) >> np.uint64(63)) & np.uint64(1)).astype(np.int16) * 2 - 1)
        return np.where(on, s16(sd, i), dither).astype(np.int16)

    return chan(seed), chan(seed ^ 0xFFFF0000)


def sweep(n, sr, f0=20.0, f1=20000.0, amp=16384.0):
    t = np.arange(n, dtype=np.float64)
    phi = 2 * np.pi * f0 * (n / sr) / np.log(f1 / f0) * ((f1 / f0) ** (t / n) - 1)
    l = np.rint(amp * np.sin(phi)).astype(np.int16)
    r = np.zeros_like(l)
    r[37:] = l[:-37]
    return l, r


def make_signal(kind, n, sr=44100, seed=1):
    if kind == "silence":
        z = np.zeros(n, dtype=np.int16)
        return z, z.copy()
    if kind == "white":
        return white(n, 0x5EED0003 + seed)
    if kind == "noise":      # moderate-level gaussian-ish noise: sum of four uniform draws
        a, b = white(n, 0x5EED0100 + seed)
        c, d = white(n, 0x5EED0200 + seed)
        return ((a.astype(np.int32) + c) >> 4).astype(np.int16), ((b.astype(np.int32) + d) >> 4).astype(np.int16)
    if kind == "octave":
        return octave_hold(n, 0x5EED0004 + seed), octave_hold(n, (0x5EED0004 + seed) ^ 0xFFFF0000)
    if kind == "burst":
        return bursts(n, 0x5EED0005 + seed)
    if kind == "sweep":
        return sweep(n, sr)
    if kind == "sine":
        t = np.arange(n)
        l = np.rint(10000 * np.sin(2 * np.pi * 440.0 * t / sr)).astype(np.int16)
        r = np.rint(8000 * np.sin(2 * np.pi * 1000.0 * t / sr)).astype(np.int16)
        return l, r
    raise ValueError(kind)
