#!/usr/bin/env python3
"""Throughput of the other BASELINE workload shapes (many independent streams per launch sequence) with PCM resident in
HBM, plus an oracle spot check of a few streams.  C2 (the bench workload) is one 10 001-frame stream; these show what the
same kernels do when the batch is wide.  usage: gpu_workloads.py [c3_streams c4_streams c5_streams]"""
import hashlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import lamejs_b200 as M
import oracle_lib as O
from synth import white, octave_hold, bursts

n3, n4, n5 = [int(x) for x in (sys.argv[1:4] + ["100", "400", "100"])[:3]] if len(sys.argv) > 1 else (100, 400, 100)
FR = 1000
N = FR * 1152

def run(name, ch, sr, kbps, streams, check):
    S = len(streams)
    flat = np.concatenate([np.concatenate(s) if ch == 2 else s[0] for s in streams])
    pcm = torch.from_numpy(flat).cuda()
    nb = M.stream_bytes(ch, sr, kbps, N)
    out = torch.zeros(S * nb + 64, dtype=torch.uint8, device="cuda")
    pcm_off = [i * ch * N for i in range(S)]; ns = [N] * S; out_off = [i * nb for i in range(S)]
    best = None
    for _ in range(3):
        tm = M.encode_streams_device(ch, sr, kbps, pcm.data_ptr(), pcm_off, ns, out.data_ptr(), out_off)
        best = tm if best is None or tm[6] < best[6] else best
    torch.cuda.synchronize()
    frames = S * M.stream_frames(N)
    audio = frames * 1152 / sr
    res = out[: S * nb].cpu().numpy()
    ok = True
    for j in check:
        ref = O.encode_stream(ch, sr, kbps, streams[j][0], streams[j][1] if ch == 2 else None)[0]
        ok &= res[j * nb:(j + 1) * nb].tobytes() == ref
    print("%s: %d streams x %d frames %dch %d Hz %d kbps: %.2f ms  = %.0f x realtime, %.2f us/frame, passes %d, oracle check of %d streams: %s"
          % (name, S, FR, ch, sr, kbps, best[6], audio / (best[6] / 1e3), 1e3 * best[6] / frames, int(best[7]), len(check), "OK" if ok else "MISMATCH"), flush=True)
    print("   kernel ms [psy, scan, mask, fb, q1, qn]:", [round(float(x), 3) for x in best[:6]], flush=True)
    return ok

ok = True
t = time.time()
s3 = [white(N, 0x5EED0003, offset=j << 32) for j in range(n3)]
ok &= run("C3 white noise", 2, 48000, 320, s3, [0, n3 // 2, n3 - 1]); del s3
s4 = [(octave_hold(N, 0x5EED0004 + 16 * j), None) for j in range(n4)]
ok &= run("C4 mono octave-hold", 1, 44100, 128, s4, [0, n4 - 1]); del s4
s5 = [bursts(N, 0x5EED0005 + 7 * j) for j in range(n5)]
ok &= run("C5 bursts (CBR)", 2, 44100, 128, s5, [0, n5 - 1]); del s5
print("total wall %.1f s" % (time.time() - t), "ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
