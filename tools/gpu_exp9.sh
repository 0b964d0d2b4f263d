#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp9_time.log
for nlb in 1 2; do
  echo "== nlb=$nlb" >> gpurun_out/exp9_time.log
  RPB200_IL_NLB=$nlb timeout 120 python tools/time_scan_il.py >> gpurun_out/exp9_time.log 2>&1
done
cat gpurun_out/exp9_time.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sort_" --csv --log-file gpurun_out/exp9_sort_launches.csv python tools/prof_kernels.py sort sortpairs > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/exp9_sort_launches.csv')) if len(r)>8]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[1:]: print(r[k][:70], r[v])
PY
