#!/bin/bash
# Bounded session: parity, timing, per-task durations of the rate loop, launch list.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== profile_run"; timeout 60 python tools/profile_run.py 10000 6 2>&1 | grep -v "^  " | tail -4
echo "== taskstat"; for k in sweep noise; do MP3B200_LIB=$PWD/lamejs_b200/libmp3b200_ts.so timeout 60 python tools/taskstat_run.py $k 2>&1 | tail -2; done
echo "== bench c2"; timeout 300 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json | cut -c1-300
bash tools/gpu_prof2.sh
