#!/bin/bash
# Round 2, call b (1 GPU): GPU parity tests at HEAD (per-stream scratch, graph replay, one-launch halo forms, SORT uniform tiles),
# the halo / sort / reduction A/Bs, ncu of the halo launches, the reference's own gtest with Base_B200 integrated, and the
# same-size incumbent timings of the kernels that changed.
TAG=${TAG:-r02_b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/time_r02.py halo sort reduce --out gpurun_out/${TAG}_time.json > gpurun_out/${TAG}_time.log 2>&1; echo "time_r02 rc=$?"
cat gpurun_out/${TAG}_time.log
timeout 200 python tools/time_r02.py halo1024 --out gpurun_out/${TAG}_time1024.json > gpurun_out/${TAG}_time1024.log 2>&1; echo "time_r02 1024 rc=$?"
cat gpurun_out/${TAG}_time1024.log
timeout 240 ncu --set full --cache-control none --clock-control none --import-source on -k regex:halo_ -c 12 -f \
    -o gpurun_out/${TAG}_halo python tools/prof_halo_r02.py > gpurun_out/${TAG}_halo_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_halo.ncu-rep --page raw --csv > gpurun_out/${TAG}_halo_raw.csv 2>/dev/null
python - <<'PY'
import csv, os
p = "gpurun_out/%s_halo_raw.csv" % os.environ.get("TAG", "r02_b")
rows = list(csv.reader(open(p)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"]
ix = [h.index(w) for w in want if w in h]
print([h[i] for i in ix], rows[1][ix[1]] if len(rows) > 1 else "")
for r in rows[2:]:
    print([r[i][:60] for i in ix])
PY
if [ -x oracle/_ref/test-raja-perf-suite-with-b200.exe ]; then
  (cd gpurun_out && timeout 600 ../oracle/_ref/test-raja-perf-suite-with-b200.exe > ${TAG}_ref_gtest.log 2>&1; echo "reference gtest rc=$?")
  grep -E "^\[|Base_B200" gpurun_out/${TAG}_ref_gtest.log | grep -E "^\[|B200" | tail -25
fi
timeout 600 python tools/ref_adjudicate.py --phases timing --kernels Stream_DOT Algorithm_REDUCE_SUM Algorithm_SORT Algorithm_SORTPAIRS \
    Apps_LTIMES Comm_HALO_PACKING_FUSED --out gpurun_out/${TAG}_adjudicate > gpurun_out/${TAG}_adjudicate.log 2>&1; echo "adjudicate rc=$?"
tail -12 gpurun_out/${TAG}_adjudicate.log
