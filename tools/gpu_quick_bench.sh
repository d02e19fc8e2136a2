#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_session.sh
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['cores'], 'kern', {k: round(v['ms'], 3) for k, v in d['kernels'].items()})"
tail -2 gpurun_out/bench.err
