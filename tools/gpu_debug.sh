#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/case.py <<'PY'
import sys, json, hashlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import lamejs_b200 as M
from synth import make_signal
G = json.load(open('tests/golden/golden.json'))
for name in sys.argv[1:]:
    g = G[name]
    l, r = make_signal(g['kind'], g['samples'], g['samplerate'], g['seed'])
    try:
        out = M.encode_streams(g['channels'], g['samplerate'], g['kbps'], [l], [r])[0]
        print(name, len(out) == g['bytes'], hashlib.sha256(out).hexdigest() == g['sha256'], flush=True)
    except Exception as e:
        print(name, "EXC", e, flush=True); break
PY
MP3B200_DEBUG_SYNC=1 timeout 100 python /tmp/case.py sine_mono_48k_256 2>&1 | tail -3
timeout 200 compute-sanitizer --tool memcheck --print-limit 3 python /tmp/case.py sine_mono_48k_256 > gpurun_out/memcheck.log 2>&1
grep -E "Invalid|at |Address|ERROR SUMMARY|True|False|EXC" gpurun_out/memcheck.log | head -20
cat > /tmp/smoke_dbg.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import lamejs_b200 as M, oracle_lib as O
from synth import make_signal
l, r = make_signal("noise", 12 * 1152 + 300, 44100, seed=7)
enc = M.Mp3Encoder(2, 44100, 128); ref = O.OracleEncoder(2, 44100, 128)
for i in range(0, len(l), 3456):
    print("call", i, flush=True)
    a = enc.encodeBuffer(l[i:i+3456], r[i:i+3456]); b = ref.encode_buffer(l[i:i+3456], r[i:i+3456])
    print(" ->", len(a), len(b), a == b, flush=True)
print("flush", flush=True)
a = enc.flush(); b = ref.flush(); print(" ->", len(a), len(b), a == b, flush=True)
PY
MP3B200_DEBUG_SYNC=1 timeout 60 python /tmp/smoke_dbg.py 2>&1 | tail -12
