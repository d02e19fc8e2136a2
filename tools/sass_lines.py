#!/usr/bin/env python3
"""Static code bytes per source line for one device function (nvdisasm --print-line-info).  usage: sass_lines.py func [lib]"""
import os, re, subprocess, sys, tempfile, collections
fn = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lamejs_b200", "libmp3b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if "config" not in f][0]
out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
cnt = collections.Counter(); inside = False; cur = None
for line in out.splitlines():
    if re.match(r"^\$\S+:", line): inside = (fn in line); continue
    if re.match(r"^\s*\.section", line): inside = False
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"^\s*/\*[0-9a-f]{4,}\*/\s+\S", line): cnt[cur] += 16
srcs = {}
for (f, l), v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    path = os.path.join(os.path.dirname(lib), "csrc", f)
    if f not in srcs and os.path.exists(path): srcs[f] = open(path).read().split("\n")
    text = srcs[f][l - 1].strip()[:100] if f in srcs else ""
    print("%6d B %-18s %5d  %s" % (v, f, l, text))
print(sum(cnt.values()), "bytes")
