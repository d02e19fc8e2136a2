#!/bin/bash
# One GPU-box session of round 2: parity tests, bench lines (all configs + reference arm), launch list of the bench command,
# full ncu captures of the main kernels, compute-sanitizer runs.  Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== profile_run"; timeout 120 python tools/profile_run.py 10000 4 2>&1 | tail -3
echo "== bench c2"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json | cut -c1-900; tail -3 gpurun_out/bench.err
if [ "$1" != "quick" ]; then
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json | cut -c1-300
for c in c3 c4 c5; do echo "== bench $c"; timeout 900 python bench.py --config $c --steps 3 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_$c.json | cut -c1-500; done
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_bench.csv
for k in k_q_outer k_subband_analysis k_psy_analysis k_q_search k_psy_masking; do
  echo "== ncu $k"
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -c 2 -f -o gpurun_out/prof_$k python tools/profile_run.py 10000 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run" gpurun_out/sanitizer_$tool.log | tail -3
done
fi
