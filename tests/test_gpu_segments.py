"""Encoder state export / import / seek and one stream cut into segments (include/mp3b200.h "encoder state",
lamejs_b200/sharding.py): the bytes always equal the single-encoder stream, i.e. the oracle's."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib  # noqa: E402
from synth import make_signal  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import lamejs_b200
    return lamejs_b200


def oracle_bytes(ch, sr, kbps, l, r):
    return oracle_lib.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)[0]


CASES = [(2, 44100, 128, "sweep", 90), (2, 44100, 128, "burst", 70), (1, 22050, 32, "noise", 80), (2, 48000, 320, "white", 40),
         (2, 44100, 128, "octave", 60)]


@pytest.mark.parametrize("ch,sr,kbps,kind,frames", CASES)
def test_checkpoint_resume_is_seamless(M, ch, sr, kbps, kind, frames):
    fs = 576 * M.granules_per_frame(ch, sr, kbps)
    n = frames * fs + 333
    l, r = make_signal(kind, n, sr, seed=5)
    want = oracle_bytes(ch, sr, kbps, l, r)
    for cut in (1, 777, 3 * fs + 224, n // 2, n - 1):
        a = M.Mp3Encoder(ch, sr, kbps)
        out = a.encodeBuffer(l[:cut], r[:cut])
        blob = a.export_state()
        a.close()
        b = M.Mp3Encoder(ch, sr, kbps)
        b.import_state(blob)
        assert b.export_state() == blob                       # a state survives the round trip unchanged
        out += b.encodeBuffer(l[cut:], r[cut:]) + b.flush()
        b.close()
        assert out == want, (kind, cut)


def test_state_of_another_configuration_is_refused(M):
    a = M.Mp3Encoder(2, 44100, 128)
    a.encodeBuffer(np.zeros(5000, np.int16), np.zeros(5000, np.int16))
    blob = a.export_state()
    b = M.Mp3Encoder(2, 44100, 192)
    with pytest.raises(Exception):
        b.import_state(blob)
    with pytest.raises(Exception):
        a.import_state(blob[:40])
    c = M.Mp3Encoder(2, 44100, 128)
    with pytest.raises(Exception):
        c.seek(3, np.zeros(10, np.int16))                      # wrong history length
    a.encodeBuffer(np.zeros(100, np.int16), np.zeros(100, np.int16))
    with pytest.raises(Exception):
        a.seek(3, np.zeros(1328, np.int16))                    # not a fresh encoder
    for e in (a, b, c):
        e.close()


@pytest.mark.parametrize("ch,sr,kbps,kind,frames", CASES)
@pytest.mark.parametrize("nseg,warmup", [(2, 8), (5, 8), (4, 1)])
def test_segments_equal_the_single_encoder(M, ch, sr, kbps, kind, frames, nseg, warmup):
    from lamejs_b200 import sharding
    fs = 576 * M.granules_per_frame(ch, sr, kbps)
    n = frames * fs + 517
    l, r = make_signal(kind, n, sr, seed=9)
    want = oracle_bytes(ch, sr, kbps, l, r)
    got, redone = sharding.encode_stream_segments_local(lambda: M.Mp3Encoder(ch, sr, kbps), l, r if ch == 2 else None, fs, nseg, warmup)
    assert got == want, (kind, nseg, warmup, redone)
    assert 0 <= redone <= nseg - 1


def test_warmup_converges_on_ordinary_material(M):
    """the speculation is useful: with 8 warm-up frames the boundary states of a long noise / sweep stream all verify"""
    from lamejs_b200 import sharding
    total = 0
    for kind in ("noise", "sweep"):
        l, r = make_signal(kind, 400 * 1152, 44100, seed=2)
        got, redone = sharding.encode_stream_segments_local(lambda: M.Mp3Encoder(2, 44100, 128), l, r, 1152, 8, 8)
        assert got == oracle_bytes(2, 44100, 128, l, r)
        total += redone
    assert total <= 2, total


def test_quiet_passage_falls_back_to_the_true_state(M):
    """ATH adjust decays over many frames in a quiet passage: a short warm-up cannot reproduce it, the hand-over catches that"""
    from lamejs_b200 import sharding
    rng = np.random.default_rng(4)
    loud = (rng.standard_normal(60 * 1152) * 6000).astype(np.int16)
    quiet = (rng.standard_normal(120 * 1152) * 12).astype(np.int16)
    l = np.concatenate([loud, quiet, loud])
    got, redone = sharding.encode_stream_segments_local(lambda: M.Mp3Encoder(1, 44100, 128), l, None, 1152, 6, 4)
    assert got == oracle_bytes(1, 44100, 128, l, None)
    print("re-encoded ranges:", redone)
