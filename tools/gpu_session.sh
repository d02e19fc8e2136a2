#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== profile_run"; timeout 120 python tools/profile_run.py 10000 3 2>&1 | tail -2
