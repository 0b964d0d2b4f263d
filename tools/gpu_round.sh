#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same command, and a
# full ncu capture of the kernels named in $1 (prof_kernels.py sections).  Outputs under gpurun_out/.
TAG=${TAG:-r01_d}
SECTIONS=${1:-"scan sort indexlist gemm"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"stream_ew|reduce_kernel|scan_tma|sort_hist_kernel|sort_onesweep_kernel<0, 0, 0>|3dpa_kernel|ltimes|halo_kernel|gemm_dmma|indexlist_tma" -c 12 -f -o gpurun_out/${TAG}_prof \
    python tools/prof_kernels.py $SECTIONS > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
