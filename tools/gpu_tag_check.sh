#!/bin/bash
# One short GPU session for the tag row: the new GPU tests, then the CRC kernel's time on a C2-sized buffer.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_gpu_tag.py -x -q > gpurun_out/tag_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tag_tests.log
tail -5 gpurun_out/tag_tests.log
timeout 60 python - > gpurun_out/tag_crc_time.json 2>gpurun_out/tag_crc_time.err <<'PY'
import json, sys, numpy as np, torch
sys.path.insert(0, "tests")
import lamejs_b200 as M
out = {}
for name, lens in [("c2_one_stream_4.18MB", [4180009]), ("c4_1000_streams_x_417KB", [417959] * 1000), ("c3_100_streams_x_960KB", [960000] * 100)]:
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    buf = torch.randint(0, 256, (int(sum(lens)) + 16,), dtype=torch.uint8, device="cuda")
    best = 1e9
    for _ in range(5):
        crc, ms = M.debug_music_crc(buf.data_ptr(), offs, lens, timed=True)
        best = min(best, ms)
    out[name] = {"bytes": int(sum(lens)), "ms": best, "GBps": sum(lens) / best / 1e6}
print(json.dumps(out))
PY
cat gpurun_out/tag_crc_time.json; tail -3 gpurun_out/tag_crc_time.err
