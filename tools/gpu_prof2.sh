#!/bin/bash
# per-kernel launch list of one C2 encode + optional full captures: tools/gpu_prof2.sh [kernel-regex ...]
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_prof.csv python tools/profile_run.py 10000 3 > gpurun_out/prof_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_prof.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
per = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1000 if u in ('ns', 'nsecond') else (v * 1000 if u in ('ms', 'msecond') else v)
    per.setdefault(r[ki].split('(')[0], []).append(v)
for k, v in per.items(): print('%-28s n=%3d  last us: %s' % (k, len(v), ' '.join('%.0f' % x for x in v[-7:])))
PY
for k in "$@"; do
  echo "== ncu $k"
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -c 2 -f -o gpurun_out/prof_$k python tools/profile_run.py 10000 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
