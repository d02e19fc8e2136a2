"""SASS of every kernel of two builds, compared instruction for instruction (addresses stripped): shows that a change left
the kernels it did not mean to touch byte-identical.  usage: python tools/sass_compare.py old.so new.so"""
import subprocess, re, hashlib, sys
def funcs(lib):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump","-sass",lib],capture_output=True,text=True).stdout
    d={}; cur=None
    for line in out.splitlines():
        m=re.match(r"\s*Function : (\S+)", line)
        if m: cur=m.group(1); d[cur]=[]; continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"/\*[0-9a-f]+\*/","",line).strip()
            d[cur].append(ins)
    return {k:(len(v),hashlib.md5("\n".join(v).encode()).hexdigest()) for k,v in d.items()}
a=funcs(sys.argv[1]); b=funcs(sys.argv[2])
for k in sorted(set(a)|set(b)):
    print(("SAME " if a.get(k)==b.get(k) else "DIFF "), k[:60], a.get(k,("-",))[0], b.get(k,("-",))[0])
