#!/bin/bash
# Last session of the round: parity on the final build, bench line, full ncu captures of the kernels not captured yet.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 300 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json | cut -c1-200
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_bench.csv
for k in k_q_finish k_q_pack k_q_prepare k_mdct k_stream_scan; do
  echo "== ncu $k"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/prof_$k python tools/profile_run.py 10000 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
