"""Derived known-answer pin for the oracle: what it emits must be a valid MPEG-1 Layer III stream that DECODES back to the
input.  tests/mp3_decode.py is an independent decoder written from ISO 11172-3 (its synthesis window comes from the decoder
table of the reference's Java tree, tests/golden/synth_window.json).  Checked properties, none of which the encoder code
can satisfy by accident: the decoded signal lines up with the input at exactly LAME's documented encoder+decoder delay of
1105 samples (576 + 529), its gain equals the preset's input scale (`Presets.js:226-244`: 0.95 up to 160 kbps, 1.0 at
320 kbps), tonal material comes back with > 60 dB SNR when bits are plentiful, and transient material that forces
START/SHORT/STOP windows still decodes coherently (short-block reorder, window shapes, subblock gains)."""
import numpy as np
import pytest

import mp3_decode
import mp3_parse
from synth import make_signal


DELAY = 1105                                             # LAME: encoder delay 576 + decoder delay 529


def _corr(y, x, d, n):
    seg, ref = y[d:d + n], x[:n]
    return float(np.dot(seg, ref) / np.sqrt(np.dot(seg, seg) * np.dot(ref, ref) + 1e-30))


def _align(y, x, n=30000):
    """correlation at the documented delay (and that it is a strict local maximum there), gain and SNR at that delay"""
    c = _corr(y, x, DELAY, n)
    sharp = c > _corr(y, x, DELAY - 1, n) and c > _corr(y, x, DELAY + 1, n)
    seg, ref = y[DELAY:DELAY + n], x[:n]
    gain = float(np.dot(seg, ref) / np.dot(ref, ref))
    snr = 10 * np.log10(np.dot(ref, ref) / np.sum((seg / gain - ref) ** 2))
    return c, sharp, gain * 32768.0, snr


@pytest.mark.parametrize("ch,sr,kbps,kind,scale,min_snr", [
    (2, 44100, 128, "sine", 0.95, 30.0), (1, 44100, 128, "sweep", 0.95, 60.0), (2, 48000, 320, "sine", 1.0, 60.0),
    (2, 32000, 160, "sweep", 0.95, 60.0), (1, 32000, 96, "sine", 0.95, 25.0), (2, 48000, 192, "sweep", 0.97, 50.0)])
def test_oracle_output_decodes_to_the_input(oracle, books, ch, sr, kbps, kind, scale, min_snr):
    l, r = make_signal(kind, 40 * 1152, sr, 3)
    data = oracle.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)[0]
    y = mp3_decode.decode(data, books)
    for c_i, x in enumerate([l, r][:ch]):
        corr, sharp, gain, snr = _align(y[c_i], x.astype(np.float64))
        assert sharp, c_i                                  # lines up at exactly 1105 samples
        assert corr > 0.999 and snr > min_snr, (c_i, corr, snr)
        assert abs(gain - scale) < 0.005, (c_i, gain)      # the preset's input scaling survives the round trip


def test_block_switching_stream_decodes(oracle, books):
    """Transient bursts at 320 kbps: the stream contains START, SHORT and STOP granules and still lines up at 1105."""
    l, r = make_signal("burst", 60 * 1152, 44100, 5)
    data = oracle.encode_stream(2, 44100, 320, l, r)[0]
    frames = mp3_parse.parse_stream(data, books)
    types = {g["block_type"] for f in frames for gr in f["gi"] for g in gr}
    assert types == {0, 1, 2, 3}
    y = mp3_decode.decode(data, books)
    corr, sharp, gain, snr = _align(y[0], l.astype(np.float64), n=50000)
    assert sharp and corr > 0.9, (corr, snr)
