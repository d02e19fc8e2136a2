#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== segments on one GPU (world 1)"; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_segments_multi.py 10000 8 2>&1 | tail -1
echo "== bench c2"; timeout 300 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
print("handle_api", d.get("handle_api")); print("e2e", d.get("e2e"))
PY
