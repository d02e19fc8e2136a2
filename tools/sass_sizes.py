#!/usr/bin/env python3
"""Static code size of every device function inside a kernel of libmp3b200.so (no GPU needed).
usage: sass_sizes.py [kernel-substring] [lib.so]"""
import os, re, subprocess, sys, tempfile
kern = sys.argv[1] if len(sys.argv) > 1 else "k_quantize_pack"
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lamejs_b200", "libmp3b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if "config" not in f][0]
out = subprocess.run(["nvdisasm", os.path.join(d, cub)], capture_output=True, text=True).stdout
sizes, cur, inside = {}, None, False
for line in out.splitlines():
    m = re.match(r"^\s*\.section\s+\.text\.(\S+?),", line)
    if m: inside = kern in m.group(1); cur = "<kernel body>" if inside else None; continue
    if not inside: continue
    m = re.match(r"^\$\S+\$(_Z\w+|\w+):", line) or re.match(r"^(\$\S+):", line)
    if m and "$_Z" in line or (m and line.startswith("$__")):
        name = m.group(1)
        dm = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        cur = dm.split("(")[0] if dm else name
        continue
    if re.match(r"^\s*/\*[0-9a-f]{4,}\*/\s+\S", line) and cur: sizes[cur] = sizes.get(cur, 0) + 16
tot = sum(sizes.values())
for k, v in sorted(sizes.items(), key=lambda kv: -kv[1]): print("%7d B  %s" % (v, k))
print("%7d B  total" % tot)
