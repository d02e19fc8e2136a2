#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run" gpurun_out/sanitizer_$tool.log | tail -3
done
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
