#!/bin/bash
# Round 2, call e (1 GPU): halo item kernel with .cg gathers, multi-variable items and unit tickets: parity, A/B, ncu, sanitizer.
TAG=${TAG:-r02_e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/time_r02.py halo --out gpurun_out/${TAG}_time.json > gpurun_out/${TAG}_time.log 2>&1; echo "time_r02 rc=$?"
cat gpurun_out/${TAG}_time.log
timeout 200 python tools/time_r02.py halo1024 --out gpurun_out/${TAG}_time1024.json > gpurun_out/${TAG}_time1024.log 2>&1; echo "time_r02 1024 rc=$?"
cat gpurun_out/${TAG}_time1024.log
timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:'halo_items_kernel|halo_kernel' -c 17 -f \
    -o gpurun_out/${TAG}_halo python tools/prof_halo_r02.py > gpurun_out/${TAG}_halo_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_halo.ncu-rep --page raw --csv > gpurun_out/${TAG}_halo_raw.csv 2>/dev/null
TAG=$TAG python - <<'PY'
import csv, os
p = "gpurun_out/%s_halo_raw.csv" % os.environ.get("TAG", "r02_e")
rows = list(csv.reader(open(p)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write_lookup_hit.sum"]
ix = [h.index(w) for w in want if w in h]
print([h[i] for i in ix]); print([rows[1][i] for i in ix])
for r in rows[2:]:
    print([r[i][:40] for i in ix])
PY
for tool in memcheck racecheck synccheck; do
  timeout 240 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py halo > gpurun_out/${TAG}_sanitize_${tool}_halo.log 2>&1
  echo "$tool halo rc=$? | $(grep -E 'SANITIZE_TARGETS|ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_${tool}_halo.log | tr '\n' ' ')"
done
