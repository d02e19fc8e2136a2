#!/bin/bash
# One GPU-box session: parity tests, timing, (on failure) first-difference report, variant A/B timing, optional ncu captures.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
if grep -q "failed" gpurun_out/pytest_gpu.log; then
  echo "== find diff"; timeout 300 python tools/gpu_find_diff.py sweep 1500 2>&1 | tail -25
fi
echo "== profile_run"; timeout 120 python tools/profile_run.py 10000 4 2>&1 | tail -3
for v in lamejs_b200/libmp3b200_*.so; do
  [ -e "$v" ] || continue
  echo "== variant $v"; MP3B200_LIB=$PWD/$v timeout 120 python tools/profile_run.py 10000 4 2>&1 | tail -3
done
for k in "$@"; do
  echo "== ncu $k"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/prof_$k python tools/profile_run.py 10000 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
