"""Minimal, independent MPEG-1 Layer III *bitstream parser* (test tool).

Parses frame header, side information, scalefactors and Huffman-coded spectral
data back to integer quantised lines (ISO/IEC 11172-3 2.4.1-2.4.3).  It is the
"decode our own bytes" pin SURVEY.md 8(c) asks for: if header bits, side info,
part2_3_length accounting, table selection, region splits, linbits or sign bits
were wrong, the parse desynchronises or the recovered lines differ from the
encoder's own l3_enc.

Huffman decode tables are rebuilt from the (code, length) pairs in the generated
constant header (tools/gen_tables.py) -- the code books themselves are ISO data.
"""
import re
import numpy as np

_BITRATES = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
_SRATES = [44100, 48000, 32000]
_SFB_L = {
    44100: [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576],
    48000: [0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576],
    32000: [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576],
}
_SFB_S = {
    44100: [0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192],
    48000: [0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192],
    32000: [0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192],
}
_SLEN1 = [0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4]
_SLEN2 = [0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3]


def _load_tables(path):
    src = open(path).read()

    def arr(name):
        m = re.search(name + r"\[\d+\] = \{(.*?)\};", src, re.S)
        return [int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()]

    return {k: arr("_" + k) for k in ["HUFF_OFF", "HUFF_XLEN", "HUFF_LINMAX", "HUFF_CODE", "HUFF_LEN"]}


class HuffBooks:
    def __init__(self, table_header):
        t = _load_tables(table_header)
        self.xlen, self.linmax = t["HUFF_XLEN"], t["HUFF_LINMAX"]
        self.dec = {}
        for tb in range(34):
            if tb in (0, 4, 14):
                continue
            n = {32: 16, 33: 16}.get(tb, None)
            side = self.xlen[tb] if tb < 16 else 16
            if n is None:
                n = side * side
            off = t["HUFF_OFF"][tb]
            d = {}
            for i in range(n):
                ln, code = t["HUFF_LEN"][off + i], t["HUFF_CODE"][off + i]
                # LAME-style length tables include the sign bits of the non-zero values
                if tb >= 32:  # count1 books: code is pre-shifted left by the number of sign bits
                    ln -= bin(i).count("1")
                    code >>= bin(i).count("1")
                else:
                    ln -= (1 if i // side else 0) + (1 if i % side else 0)
                key = (ln, code)
                assert key not in d, ("ambiguous code", tb, i)
                d[key] = i
            self.dec[tb] = (d, side, max(k[0] for k in d))

    def decode(self, br, tb):
        d, side, maxlen = self.dec[tb]
        code = 0
        for ln in range(1, maxlen + 1):
            code = (code << 1) | br.get(1)
            v = d.get((ln, code))
            if v is not None:
                return v
        raise ValueError("bad huffman code in table %d" % tb)


class BitReader:
    def __init__(self, data, pos=0):
        self.d, self.p = data, pos

    def get(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | ((self.d[self.p >> 3] >> (7 - (self.p & 7))) & 1)
            self.p += 1
        return v


def _parse_main(br, gi, scfsi, nch, sr, books):
    """scalefactors and Huffman data of one frame's granules from `br` (positioned at the frame's main data)"""
    sfl, sfs = _SFB_L[sr], _SFB_S[sr]
    for gr in range(2):
        for ch in range(nch):
            g = gi[gr][ch]
            start = br.p
            s1, s2 = _SLEN1[g["scalefac_compress"]], _SLEN2[g["scalefac_compress"]]
            sf = []
            if g["block_type"] == 2:
                for sfb in range(12):
                    for w in range(3):
                        sf.append(br.get(s1 if sfb < 6 else s2))
            else:
                for sfb in range(21):
                    band = 0 if sfb < 6 else 1 if sfb < 11 else 2 if sfb < 16 else 3
                    if gr == 1 and scfsi[ch][band]:
                        sf.append(gi[0][ch]["scalefac"][sfb])
                    else:
                        sf.append(br.get(s1 if sfb < 11 else s2))
            g["scalefac"] = sf
            g["part2_length"] = br.p - start
            end = start + g["part2_3_length"]
            ix = np.zeros(576, dtype=np.int32)
            bv = g["big_values"]
            if g["block_type"] == 2:
                r1, r2 = min(3 * sfs[3], bv), bv
            elif g["window_switching"]:
                r1, r2 = min(sfl[8], bv), bv
            else:
                r1 = min(sfl[g["region0"] + 1], bv)
                r2 = min(sfl[g["region0"] + g["region1"] + 2], bv)
            i = 0
            while i < bv:
                tb = g["table_select"][0 if i < r1 else 1 if i < r2 else 2]
                if tb == 0:
                    x = y = 0
                else:
                    v = books.decode(br, tb)
                    side = books.dec[tb][1]
                    x, y = v // side, v % side
                    lin = books.xlen[tb] if tb > 15 else 0
                    if tb > 15 and x == 15:
                        x += br.get(lin)
                    if x and br.get(1):
                        x = -x
                    if tb > 15 and y == 15:
                        y += br.get(lin)
                    if y and br.get(1):
                        y = -y
                ix[i], ix[i + 1] = x, y
                i += 2
            tb = 32 + g["count1table"]
            while br.p < end and i <= 572:
                v = books.decode(br, tb)
                q = [(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1]
                for k in range(4):
                    if q[k] and br.get(1):
                        q[k] = -1
                ix[i:i + 4] = q
                i += 4
            assert br.p == end, ("part2_3_length mismatch", gr, ch, br.p - end)
            g["ix"], g["count1_end"] = ix, i


def parse_frame(data, off, books, main=None):
    """Parse one frame starting at byte `off`.  Returns dict with header, side info, per granule/channel
    quantised lines `ix` (signed), scalefactors, bit accounting and frame length."""
    br = BitReader(data, off * 8)
    assert br.get(12) == 0xFFF, "sync"
    assert br.get(1) == 1 and br.get(2) == 1, "MPEG-1 layer III"
    prot = br.get(1)
    bri, sri, pad = br.get(4), br.get(2), br.get(1)
    br.get(1)
    mode, mode_ext = br.get(2), br.get(2)
    br.get(4)
    sr = _SRATES[sri]
    nch = 1 if mode == 3 else 2
    flen = 144000 * _BITRATES[bri] // sr + pad
    mdb = br.get(9)
    br.get(5 if nch == 1 else 3)
    scfsi = [[br.get(1) for _ in range(4)] for _ in range(nch)]
    gi = [[None] * nch for _ in range(2)]
    for gr in range(2):
        for ch in range(nch):
            g = {"part2_3_length": br.get(12), "big_values": br.get(9) * 2, "global_gain": br.get(8),
                 "scalefac_compress": br.get(4), "window_switching": br.get(1)}
            if g["window_switching"]:
                g["block_type"], g["mixed"] = br.get(2), br.get(1)
                g["table_select"] = [br.get(5), br.get(5), 0]
                g["subblock_gain"] = [br.get(3), br.get(3), br.get(3)]
                g["region0"], g["region1"] = (8 if g["block_type"] == 2 else 7), 36
            else:
                g["block_type"], g["mixed"] = 0, 0
                g["table_select"] = [br.get(5), br.get(5), br.get(5)]
                g["subblock_gain"] = [0, 0, 0]
                g["region0"], g["region1"] = br.get(4), br.get(3)
            g["preflag"], g["scalefac_scale"], g["count1table"] = br.get(1), br.get(1), br.get(1)
            gi[gr][ch] = g
    side_end = br.p
    assert side_end == (off + 4 + (17 if nch == 1 else 32)) * 8
    if main is None:
        assert mdb == 0, "bit reservoir not expected"
        _parse_main(br, gi, scfsi, nch, sr, books)
        main_end = br.p - off * 8
    else:
        # bit reservoir: this frame's main data starts `mdb` bytes before the end of what earlier frames carried
        start = len(main) - mdb
        assert start >= 0, "main_data_begin points before the stream"
        main += data[off + 4 + (17 if nch == 1 else 32):off + flen]
        mbr = BitReader(main, start * 8)
        _parse_main(mbr, gi, scfsi, nch, sr, books)
        assert mbr.p <= len(main) * 8, "main data runs past this frame"
        main_end = None
    return {"nch": nch, "sr": sr, "kbps": _BITRATES[bri], "padding": pad, "frame_len": flen, "mode": mode,
            "mode_ext": mode_ext, "scfsi": scfsi, "gi": gi, "main_end_bit": main_end, "main_data_begin": mdb}


def parse_stream(data, books):
    off, frames = 0, []
    while off + 4 <= len(data):
        f = parse_frame(data, off, books)
        frames.append(f)
        off += f["frame_len"]
    assert off == len(data), "trailing bytes"
    return frames


def parse_stream_reservoir(data, books):
    """Like parse_stream for streams written with the bit reservoir: frames sit on the fixed grid, each frame's main data
    starts main_data_begin bytes before its own side info ends (ISO 11172-3 2.4.2.7), gathered here in one byte string."""
    off, frames, main = 0, [], bytearray()
    while off + 4 <= len(data):
        f = parse_frame(data, off, books, main)
        frames.append(f)
        off += f["frame_len"]
    assert off == len(data), "trailing bytes"
    return frames
