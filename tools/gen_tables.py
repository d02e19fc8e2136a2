#!/usr/bin/env python3
"""Development-time generator for the constant data tables of the MP3 path.

The MPEG-1 Layer III constants (ISO 11172-3 Huffman code books, scalefactor
band edges, the Takehiro-folded 512-tap analysis window, MDCT windows, the
FFT bit-reversal table) are *data*.  lamejs spells several of them as
arithmetic expressions evaluated by the JS engine in IEEE double, left to
right (reference: src/js/NewMDCT.js:52-337, :342-510, src/js/Tables.js:11-507,
src/js/FFT.js:24-29,117-138).  This tool evaluates those expressions with
Python floats (also IEEE double, same left-to-right order) and emits them in
this repo's own flat layout as hex-float / integer C arrays, once for the
oracle (oracle/lj_tables.h) and once for the product (lamejs_b200/csrc/
mp3_tables.h).  It needs /root/reference and is therefore run only in the
build container; the generated headers are committed.

Layout of the generated Huffman data (ours, not the reference's):
  HUFF_OFF[t]   start of code book t (0..33) inside HUFF_CODE / HUFF_LEN
  HUFF_XLEN[t]  row stride (ht[t].xlen), HUFF_LINMAX[t] = ht[t].linmax
"""
import re
import sys

REF = "/root/reference/src/js/"


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = re.sub(r"//[^\n]*", "", s)
    return s


def grab_array(src, header_regex):
    """Return the python-evaluated value of the JS array literal that follows
    header_regex."""
    m = re.search(header_regex, src)
    if not m:
        raise KeyError(header_regex)
    i = src.index("[", m.end() - 1)
    depth = 0
    j = i
    while True:
        c = src[j]
        if c == "[":
            depth += 1
        elif c == "]":
            depth -= 1
            if depth == 0:
                break
        j += 1
    body = src[i : j + 1]
    body = body.replace("Util.SQRT2", "1.41421356237309504880")
    return eval(body, {"__builtins__": {}}, {})


def hexf(x):
    return float(x).hex()


def emit_f64(name, vals, per=4):
    out = ["static const double %s[%d] = {" % (name, len(vals))]
    for k in range(0, len(vals), per):
        out.append("  " + ", ".join(hexf(v) for v in vals[k : k + per]) + ",")
    out.append("};")
    return "\n".join(out)


def emit_int(name, vals, ctype="int", per=16):
    out = ["static const %s %s[%d] = {" % (ctype, name, len(vals))]
    for k in range(0, len(vals), per):
        out.append("  " + ", ".join(str(int(v)) for v in vals[k : k + per]) + ",")
    out.append("};")
    return "\n".join(out)


def build(prefix):
    P = prefix
    mdct = strip_comments(open(REF + "NewMDCT.js").read())
    tabs = strip_comments(open(REF + "Tables.js").read())
    fft = strip_comments(open(REF + "FFT.js").read())

    parts = []
    enwindow = grab_array(mdct, r"var enwindow = \[")
    win = grab_array(mdct, r"var win = \[")
    order = grab_array(mdct, r"var order = \[")
    assert len(win) == 4 and all(len(w) == 36 for w in win)
    parts.append(emit_f64(P + "ENWINDOW", enwindow))
    parts.append(emit_f64(P + "MDCT_WIN", [v for w in win for v in w]))
    parts.append(emit_int(P + "SB_ORDER", order))

    costab = grab_array(fft, r"var costab = \[")
    rv = grab_array(fft, r"var rv_tbl = \[")
    parts.append(emit_f64(P + "FHT_COSTAB", costab))
    parts.append(emit_int(P + "FFT_RV", rv, "unsigned char"))

    # Huffman code books -> flat layout
    ht_meta = re.findall(
        r"new HuffCodeTab\((\d+),\s*(\d+),\s*(null|Tables\.\w+),\s*(null|Tables\.\w+)\)",
        tabs,
    )
    assert len(ht_meta) == 34
    codes, lens, off, xlen, linmax = [], [], [], [], []
    seen = {}
    for (xl, lm, tb, hl) in ht_meta:
        xlen.append(int(xl))
        linmax.append(int(lm))
        if (tb, hl) in seen:          # books 16..23 / 24..31 share one code table
            off.append(seen[(tb, hl)])
            continue
        off.append(len(codes))
        seen[(tb, hl)] = len(codes)
        if hl == "null":
            continue
        L = grab_array(tabs, re.escape(hl) + r" = \[")
        C = [0] * len(L) if tb == "null" else grab_array(tabs, re.escape(tb) + r" = \[")
        assert len(C) == len(L), (tb, hl)
        codes += C
        lens += L
    parts.append(emit_int(P + "HUFF_OFF", off))
    parts.append(emit_int(P + "HUFF_XLEN", xlen))
    parts.append(emit_int(P + "HUFF_LINMAX", linmax))
    parts.append(emit_int(P + "HUFF_CODE", codes, "unsigned short"))
    parts.append(emit_int(P + "HUFF_LEN", lens, "unsigned char"))
    parts.append(emit_int(P + "HUFF_LARGETBL", grab_array(tabs, r"Tables\.largetbl = \["), "unsigned int", 8))
    parts.append(emit_int(P + "HUFF_TABLE23", grab_array(tabs, r"Tables\.table23 = \["), "unsigned int", 8))
    parts.append(emit_int(P + "HUFF_TABLE56", grab_array(tabs, r"Tables\.table56 = \["), "unsigned int", 8))
    return "\n\n".join(parts) + "\n"


def main():
    targets = [
        ("oracle/lj_tables.h", "LJ_", "LJ_TABLES_H"),
        ("lamejs_b200/csrc/mp3_tables.h", "MP3_", "MP3B200_TABLES_H"),
    ]
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/repo/"
    for path, prefix, guard in targets:
        body = build(prefix)
        with open(root + path, "w") as f:
            f.write("/* GENERATED by tools/gen_tables.py -- ISO 11172-3 / LAME constant data, own layout. */\n")
            f.write("#ifndef %s\n#define %s\n\n" % (guard, guard))
            f.write(body)
            f.write("\n#endif\n")
        print("wrote", path)


if __name__ == "__main__":
    main()
