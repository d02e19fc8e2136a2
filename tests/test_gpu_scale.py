"""BASELINE.json configs #3/#4/#5 at (or near) their stated sizes, EVERY stream compared byte for byte with the CPU oracle
(thread pool over the host cores; ctypes releases the GIL), plus a randomised soak of >= 30 k frames.  The speculate /
verify / re-validate machinery of the quantizer (cross-frame OldValue recurrence) is exactly the kind of logic whose
rare failure only shows up at scale."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from synth import bursts, make_signal, octave_hold, white

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import lamejs_b200 as m
    return m


def _check_all(M, oracle, ch, sr, kbps, ls, rs):
    oracle.lib()
    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        futs = [ex.submit(lambda j=j: oracle.encode_stream(ch, sr, kbps, ls[j], rs[j] if ch == 2 else None)[0]) for j in range(len(ls))]
        outs = M.encode_streams(ch, sr, kbps, ls, rs if ch == 2 else None)
        bad = [j for j, f in enumerate(futs) if outs[j] != f.result()]
    assert not bad, "streams differing from the oracle: %s" % bad[:10]
    return outs


def test_c3_full_size_white_noise_320k(M, oracle):
    """config #3: stereo 48 kHz 320 kbps, 100 streams x 1000 white-noise frames (stream j uses counter offset j * 2^32)."""
    S, n = 100, 1000 * 1152
    ls, rs = zip(*[white(n, 0x5EED0003, offset=j << 32) for j in range(S)])
    outs = _check_all(M, oracle, 2, 48000, 320, list(ls), list(rs))
    assert all(len(o) == 1001 * 960 for o in outs)


def test_c4_mono_octave_200_streams(M, oracle):
    """config #4 shape: mono 44.1 kHz 128 kbps octave-hold noise, 200 streams x 1000 frames (the 1000-stream run is a bench config)."""
    S, n = 200, 1000 * 1152
    ls = [octave_hold(n, 0x5EED0004 + 16 * j) for j in range(S)]
    _check_all(M, oracle, 1, 44100, 128, ls, ls)


def test_c5_full_size_bursts(M, oracle):
    """config #5 input (transient bursts -> START/SHORT/STOP switching) under CBR 128k: 100 streams x 1000 frames."""
    S, n = 100, 1000 * 1152
    ls, rs = zip(*[bursts(n, 0x5EED0005 + 64 * j) for j in range(S)])
    _check_all(M, oracle, 2, 44100, 128, list(ls), list(rs))


def test_random_soak_30k_frames(M, oracle):
    rng = np.random.default_rng(20260924)
    kinds = ["noise", "burst", "sweep", "white", "sine", "octave", "silence"]
    configs = [(ch, sr, kbps) for sr in (32000, 44100, 48000) for kbps in (64, 96, 112, 128, 160, 192, 256, 320) for ch in (1, 2)
               if M.stream_bytes(ch, sr, kbps, 1152) > 0]
    by_cfg, frames = {}, 0
    for i in range(150):
        cfg = configs[rng.integers(len(configs))]
        kind = kinds[rng.integers(len(kinds))]
        n = int(rng.integers(1, 420 * 1152))
        by_cfg.setdefault(cfg, []).append(make_signal(kind, n, cfg[1], int(rng.integers(1 << 30))))
        frames += M.stream_frames(n)
    assert frames >= 30000
    for (ch, sr, kbps), sigs in by_cfg.items():
        _check_all(M, oracle, ch, sr, kbps, [s[0] for s in sigs], [s[1] for s in sigs])
