#!/bin/bash
# standard GPU validation session; each step under its own timeout, logs in gpurun_out/
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== profile_run"; timeout 120 python tools/profile_run.py 10000 3 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01.json | cut -c1-1200
