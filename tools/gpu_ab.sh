#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do echo "== profile_run"; timeout 60 python tools/profile_run.py 10000 6 2>&1 | grep -v "^  " | tail -4; done
bash tools/gpu_prof2.sh
