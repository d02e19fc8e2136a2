"""A small MPEG-1 Layer III decoder written from ISO/IEC 11172-3 (requantisation 2.4.3.4, alias reduction, IMDCT with
the four block-type windows, frequency inversion, polyphase synthesis) on top of the bitstream parser in mp3_parse.py.
Independent of the encoder under test: it shares no code and no derivation with lamejs, the oracle or the CUDA kernels
(the synthesis prototype window comes from the decoder table in the reference's Java tree, see
tools/gen_decoder_window.py).  Stereo modes: L/R and M/S (lamejs' Mp3Encoder never sets mode_ext; the oracle's joint-stereo mode does).  numpy, slow,
meant for a few dozen frames in tests: decoding the oracle's output back to the input is a derived known-answer pin."""
import json
import os

import numpy as np

import mp3_parse

_HERE = os.path.dirname(os.path.abspath(__file__))
_PRETAB = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0]
_CI = np.array([-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037])
_CS = 1.0 / np.sqrt(1.0 + _CI * _CI)
_CA = _CI / np.sqrt(1.0 + _CI * _CI)


def _synthesis_window():
    dewin = np.array(json.load(open(os.path.join(_HERE, "golden", "synth_window.json")))["dewin"])
    i = np.arange(512)
    return ((-1.0) ** (i // 64)) * dewin[np.minimum(i, 512 - i)]       # ISO table D[i]


def _imdct_windows():
    i = np.arange(36)
    w = np.zeros((4, 36))
    w[0] = np.sin(np.pi / 36 * (i + 0.5))
    w[1, :18] = w[0, :18]; w[1, 18:24] = 1.0; w[1, 24:30] = np.sin(np.pi / 12 * (np.arange(24, 30) - 18 + 0.5))
    w[3, 6:12] = np.sin(np.pi / 12 * (np.arange(6, 12) - 6 + 0.5)); w[3, 12:18] = 1.0; w[3, 18:] = w[0, 18:]
    return w, np.sin(np.pi / 12 * (np.arange(12) + 0.5))


class Decoder:
    def __init__(self, nch, sr):
        self.nch, self.sr = nch, sr
        self.D = _synthesis_window()
        self.win, self.win_s = _imdct_windows()
        i, k = np.arange(36)[:, None], np.arange(18)[None, :]
        self.cos36 = np.cos(np.pi / 72 * (2 * i + 1 + 18) * (2 * k + 1))          # 36 x 18
        i, k = np.arange(12)[:, None], np.arange(6)[None, :]
        self.cos12 = np.cos(np.pi / 24 * (2 * i + 1 + 6) * (2 * k + 1))           # 12 x 6
        i, k = np.arange(64)[:, None], np.arange(32)[None, :]
        self.N = np.cos((16 + i) * (2 * k + 1) * np.pi / 64)                      # 64 x 32
        self.prev = np.zeros((nch, 32, 18))
        self.V = np.zeros((nch, 1024))
        self.sfl, self.sfs = mp3_parse._SFB_L[sr], mp3_parse._SFB_S[sr]

    def _requantize(self, g):
        ix = g["ix"].astype(np.float64)
        mag = np.sign(ix) * np.abs(ix) ** (4.0 / 3.0)
        mult = 1.0 if g["scalefac_scale"] else 0.5
        gg = g["global_gain"]
        xr = np.zeros(576)
        sf = g["scalefac"]
        if g["block_type"] == 2:
            per_window = np.zeros((3, 192))
            for sfb in range(13):
                lo, hi = self.sfs[sfb], self.sfs[sfb + 1]
                wd = hi - lo
                for w in range(3):
                    s = sf[3 * sfb + w] if sfb < 12 else 0
                    scale = 2.0 ** ((gg - 210 - 8 * g["subblock_gain"][w]) / 4.0) * 2.0 ** (-(mult * s))
                    pos = 3 * lo + w * wd
                    per_window[w, lo:hi] = mag[pos:pos + wd] * scale
            return None, per_window
        for sfb in range(22):
            lo, hi = self.sfl[sfb], self.sfl[sfb + 1]
            s = (sf[sfb] + (g["preflag"] * _PRETAB[sfb])) if sfb < 21 else 0
            xr[lo:hi] = mag[lo:hi] * 2.0 ** ((gg - 210) / 4.0) * 2.0 ** (-(mult * s))
        return xr, None

    def _hybrid(self, ch, g, xr, per_window):
        out = np.zeros((18, 32))                    # [time slot][subband]
        bt = g["block_type"]
        if bt != 2:
            x = xr.copy()
            for sb in range(1, 32):                 # alias reduction butterflies
                for i in range(8):
                    lo, hi = 18 * sb - 1 - i, 18 * sb + i
                    a, b = x[lo], x[hi]
                    x[lo] = a * _CS[i] - b * _CA[i]
                    x[hi] = b * _CS[i] + a * _CA[i]
        for sb in range(32):
            if bt == 2:
                res = np.zeros(36)
                for w in range(3):
                    y = self.cos12 @ per_window[w, 6 * sb:6 * sb + 6] * self.win_s
                    res[6 + 6 * w:18 + 6 * w] += y
            else:
                res = (self.cos36 @ x[18 * sb:18 * sb + 18]) * self.win[bt]
            t = res[:18] + self.prev[ch, sb]
            self.prev[ch, sb] = res[18:]
            if sb & 1:
                t[1::2] = -t[1::2]                  # frequency inversion
            out[:, sb] = t
        return out

    def _polyphase(self, ch, slots):
        pcm = np.zeros(18 * 32)
        V = self.V[ch]
        for t in range(18):
            V[64:] = V[:-64].copy()
            V[:64] = self.N @ slots[t]
            U = np.empty(512)
            for a in range(8):
                U[64 * a:64 * a + 32] = V[128 * a:128 * a + 32]
                U[64 * a + 32:64 * a + 64] = V[128 * a + 96:128 * a + 128]
            pcm[32 * t:32 * t + 32] = (U * self.D).reshape(16, 32).sum(0)
        return pcm

    def decode_frame(self, f):
        out = np.zeros((self.nch, 1152))
        ms = self.nch == 2 and f["mode"] == 1 and (f["mode_ext"] & 2) != 0     # joint stereo, M/S on (ISO 11172-3 2.4.3.4.9.3)
        for gr in range(2):
            rq = [self._requantize(f["gi"][gr][ch]) for ch in range(self.nch)]
            if ms:
                # both channels carry the same block type in an M/S granule: L = (M + S) / sqrt 2, R = (M - S) / sqrt 2
                k = 0 if rq[0][0] is not None else 1
                m, s_ = rq[0][k], rq[1][k]
                lr = ((m + s_) / np.sqrt(2.0), (m - s_) / np.sqrt(2.0))
                rq = [(lr[c], None) if k == 0 else (None, lr[c]) for c in range(2)]
            for ch in range(self.nch):
                g = f["gi"][gr][ch]
                out[ch, 576 * gr:576 * gr + 576] = self._polyphase(ch, self._hybrid(ch, g, rq[ch][0], rq[ch][1]))
        return out


def decode(data, books, max_frames=None, reservoir=False):
    """bytes -> float array [nch][frames * 1152] (full scale = 32768, like the encoder's Int16 input)."""
    frames = mp3_parse.parse_stream_reservoir(data, books) if reservoir else mp3_parse.parse_stream(data, books)
    if max_frames:
        frames = frames[:max_frames]
    dec = Decoder(frames[0]["nch"], frames[0]["sr"])
    return np.concatenate([dec.decode_frame(f) for f in frames], axis=1)
