#!/bin/bash
# Bounded session: parity, timing A/B (MDCT beside / behind the masking kernel), launch list.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do
echo "== profile_run"; timeout 60 python tools/profile_run.py 10000 6 2>&1 | grep -v "^  " | tail -4
echo "== profile_run serial mdct"; MP3B200_SERIAL_MDCT=1 timeout 60 python tools/profile_run.py 10000 6 2>&1 | grep -v "^  " | tail -4
done
echo "== bench c2"; timeout 300 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json | cut -c1-300
echo "== bench c4"; timeout 300 python bench.py --config c4 --steps 3 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_c4.json | cut -c1-300
bash tools/gpu_prof2.sh
