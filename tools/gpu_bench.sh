#!/bin/bash
# Round-end measurement on the GPU box: bench line, reference arm, ncu launch list of the same bench command, one
# --set full capture of each kernel.  Outputs under gpurun_out/ (copied into profiles/ by hand).
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json | cut -c1-400
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/launches.csv
for k in "$@"; do
  echo "== ncu $k"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/prof_$k python tools/profile_run.py 10000 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
