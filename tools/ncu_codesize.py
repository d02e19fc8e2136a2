#!/usr/bin/env python3
"""Code bytes and executed-instruction share per device function from an `ncu --page source --csv --print-source cuda,sass`
export (functions located by line range in the given source file).  usage: ncu_codesize.py export.csv k_quant.cuh"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
srcfile = sys.argv[2]
funcs = []
for n, line in enumerate(open(srcfile), 1):
    m = re.match(r"^__device__ \w+ .*?(\w+)\(", line)
    if m: funcs.append((n, m.group(1)))
    if line.startswith("k_quantize_pack("): funcs.append((n - 1, "k_quantize_pack"))
def fo(ln):
    nm = "?"
    for s, f in funcs:
        if s <= ln: nm = f
    return nm
cur = hdr = curline = None
size = collections.Counter(); ie = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No":
        hdr = {}
        for i, k in enumerate(r): hdr.setdefault(k, i)
        continue
    if hdr is None: continue
    if r[0] != "":
        try: curline = int(r[0])
        except ValueError: pass
        continue
    if not r[2].startswith("0x"): continue
    key = fo(curline) if cur.endswith(srcfile.split("/")[-1]) else "[" + cur.split("/")[-1] + "]"
    size[key] += 16; ie[key] += int(r[hdr["Instructions Executed"]] or 0)
ti = sum(ie.values()); cum = 0
for k, v in sorted(size.items(), key=lambda kv: -ie[kv[0]]):
    cum += v
    print("%-28s %6d B  ins %5.2f%%  cum %7d B" % (k, v, 100 * ie[k] / ti, cum))
print("total", sum(size.values()))
