#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== profile_run"; timeout 120 python tools/profile_run.py 10000 3 2>&1 | tail -3
