#!/bin/bash
# Round 2, call d (1 GPU): as call c after the unit-count fix (tests/test_halo_units_cpu.py now checks the unit lists on the CPU).
TAG=${TAG:-r02_d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/time_r02.py halo sort --out gpurun_out/${TAG}_time.json > gpurun_out/${TAG}_time.log 2>&1; echo "time_r02 rc=$?"
cat gpurun_out/${TAG}_time.log
[ -x tools/bin/incumbent ] && (timeout 120 tools/bin/incumbent > gpurun_out/${TAG}_cub.jsonl 2>&1; cat gpurun_out/${TAG}_cub.jsonl)
timeout 200 python tools/time_r02.py halo1024 --out gpurun_out/${TAG}_time1024.json > gpurun_out/${TAG}_time1024.log 2>&1; echo "time_r02 1024 rc=$?"
cat gpurun_out/${TAG}_time1024.log
timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:'halo_items_kernel|halo_kernel' -c 17 -f \
    -o gpurun_out/${TAG}_halo python tools/prof_halo_r02.py > gpurun_out/${TAG}_halo_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_halo.ncu-rep --page raw --csv > gpurun_out/${TAG}_halo_raw.csv 2>/dev/null
TAG=$TAG python - <<'PY'
import csv, os
p = "gpurun_out/%s_halo_raw.csv" % os.environ.get("TAG", "r02_d")
rows = list(csv.reader(open(p)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
ix = [h.index(w) for w in want if w in h]
print([h[i] for i in ix]); print([rows[1][i] for i in ix])
for r in rows[2:]:
    print([r[i][:40] for i in ix])
PY
timeout 480 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json | head -60
timeout 600 python tools/ref_adjudicate.py --phases timing --kernels Algorithm_SORT Algorithm_SORTPAIRS \
    Comm_HALO_PACKING_FUSED --out gpurun_out/${TAG}_adjudicate > gpurun_out/${TAG}_adjudicate.log 2>&1; echo "adjudicate rc=$?"
tail -8 gpurun_out/${TAG}_adjudicate.log
