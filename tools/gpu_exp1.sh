#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_comm_gpu.py tests/test_algorithm_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/time_quick.py halo sort > gpurun_out/exp1_time.log 2>&1; tail -30 gpurun_out/exp1_time.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sort_onesweep" -c 4 -f -o gpurun_out/exp1_sort python tools/prof_kernels.py sort sortpairs > gpurun_out/exp1_prof.log 2>&1; echo "ncu rc=$?"
