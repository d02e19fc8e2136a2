#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: samples / instructions / stall reasons per
function (by line range of k_quant.cuh) and per source line.   usage: ncu_lines.py export.csv k_quant.cuh [topN]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
srcfile = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
funcs = []
for n, line in enumerate(open(srcfile), 1):
    m = re.match(r"^(?:__device__|__global__|static)[^(]*?(\w+)\s*\($", line.split("(")[0] + "(") if "(" in line else None
    if m and (line.startswith("__device__") or line.startswith("__global__")):
        funcs.append((n, m.group(1)))
    elif line.startswith("k_quantize_pack("):
        funcs.append((n - 1, "k_quantize_pack"))
funcs = [f for f in funcs if f[1] not in ("__launch_bounds__",)]
def func_of(ln):
    name = "?"
    for s, f in funcs:
        if s <= ln: name = f
    return name
def I(x):
    try: return int(x)
    except Exception: return 0
cols = ["stall_barrier", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving"]
perf = collections.defaultdict(collections.Counter)
perl = collections.defaultdict(collections.Counter)
srcs = {}
cur_file, hdr = None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Line No":
        first = {}
        for i, k in enumerate(r): first.setdefault(k, i)
        hdr = first; continue
    if hdr is None: continue
    try: ln = int(r[hdr["Line No"]])
    except Exception: continue
    base = cur_file.split("/")[-1]
    key = func_of(ln) if cur_file.endswith(srcfile.split("/")[-1]) else "[" + base + "]"
    c = perf[key]; l = perl[(base, ln)]
    for cc in (c, l):
        cc["smp"] += I(r[hdr["# Samples"]]); cc["ie"] += I(r[hdr["Instructions Executed"]])
        for k in cols: cc[k] += I(r[hdr[k]])
    srcs[(base, ln)] = r[1]
ts = sum(c["smp"] for c in perf.values()); ti = sum(c["ie"] for c in perf.values())
print("total samples %d, warp instructions %d" % (ts, ti))
print("%-26s %7s %7s | %s" % ("function", "smp%", "ins%", " ".join(k[6:12].rjust(7) for k in cols)))
for f, c in sorted(perf.items(), key=lambda kv: -kv[1]["smp"]):
    print("%-26s %6.2f%% %6.2f%% | %s" % (f, 100 * c["smp"] / ts, 100 * c["ie"] / ti, " ".join(("%6.1f%%" % (100 * c[k] / max(1, c["smp"]))) for k in cols)))
print()
for (b, ln), c in sorted(perl.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    print("%-16s %5d smp %5.2f%% ins %5.2f%% noinst %5d  %s" % (b[:16], ln, 100 * c["smp"] / ts, 100 * c["ie"] / ti, c["stall_no_inst"], srcs[(b, ln)][:90]))
