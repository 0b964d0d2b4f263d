#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same command, and a
# full ncu capture of the kernels named in $1 (prof_kernels.py sections).  Outputs under gpurun_out/.
TAG=${TAG:-r01_e}
SECTIONS=${1:-"scan indexlist gemm pa"}
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
    -k regex:"stream_ew|reduce_kernel|scan_tma|3dpa_kernel|ltimes|halo_kernel|gemm_dmma|indexlist_tma" -c 12 -f -o gpurun_out/${TAG}_prof \
    python tools/prof_kernels.py $SECTIONS > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
# the C++ suite harness at the BASELINE sizes: every kernel, --graph, one bandwidth report
H=rajaperf_b200/suite/raja-perf-b200.exe
O=gpurun_out/${TAG}_harness; rm -rf $O; mkdir -p $O
( timeout 150 $H -k Stream MEMCPY MEMSET --size 268435456 --checkrun 20 --graph --outdir $O/stream
  timeout 150 $H -k REDUCE_SUM SCAN INDEXLIST INDEXLIST_3LOOP --size 134217728 --checkrun 20 --graph --outdir $O/algo
  timeout 150 $H -k SORT SORTPAIRS --size 134217728 --checkrun 2 --graph --outdir $O/sort
  timeout 150 $H -k MASS3DPA --size 500000000 --checkrun 10 --graph --outdir $O/mass
  timeout 150 $H -k DIFFUSION3DPA CONVECTION3DPA --size 256000000 --checkrun 10 --graph --outdir $O/pa
  timeout 150 $H -k LTIMES --size 1024000000 --checkrun 10 --graph --outdir $O/ltimes
  timeout 150 $H -k Polybench_GEMM --size 16777216 --checkrun 5 --graph --outdir $O/gemm
  timeout 150 $H -k HALO_PACKING HALO_PACKING_FUSED HALO_SENDRECV HALO_EXCHANGE HALO_EXCHANGE_FUSED --size 134217728 --checkrun 50 --graph --outdir $O/comm ) > gpurun_out/${TAG}_harness.log 2>&1
( head -1 $O/stream/RAJAPerf-bandwidth.csv; for d in stream algo sort mass pa ltimes gemm comm; do tail -n +2 $O/$d/RAJAPerf-bandwidth.csv; done ) > gpurun_out/${TAG}_harness_bandwidth.csv
cat gpurun_out/${TAG}_harness_bandwidth.csv | cut -d, -f1,3,4,8,9,12
ls -la gpurun_out | head -30
