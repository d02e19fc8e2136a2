#!/usr/bin/env python3
"""Per-task durations of the rate loop (Q_TASKSTAT build): one C2 encode, dumps gpurun_out/taskstat.npy
columns: clocks>>6, gr, max_nonzero_coeff, block type, gain in, bits in, target bits, gain out (row = granule-channel)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import lamejs_b200 as M  # noqa: E402
from synth import make_signal  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "sweep"
frames = 10000
n = frames * 1152
l, r = make_signal(kind, n, 44100)
pcm = torch.from_numpy(np.concatenate([l, r])).cuda()
nb = M.stream_bytes(2, 44100, 128, n)
out = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    tm = M.encode_streams_device(2, 44100, 128, pcm.data_ptr(), [0], [n], out.data_ptr(), [0])
torch.cuda.synchronize()
rows = 2 * 2 * (frames + 1)
buf = np.zeros((rows, 8), dtype=np.int32)
rc = M.lib().mp3b200_debug_taskstat(buf.ctypes.data_as(ctypes.c_void_p), rows)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "taskstat_%s.npy" % kind), buf)
d = buf[:, 0].astype(float)
print("rc", rc, "tasks", (d > 0).sum(), "mean", d[d > 0].mean(), "p50", np.percentile(d[d > 0], 50), "p99", np.percentile(d[d > 0], 99), "max", d.max())
