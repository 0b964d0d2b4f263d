#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widened_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/time_quick.py indexlist gemm > gpurun_out/exp5_time.log 2>&1; cat gpurun_out/exp5_time.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_dmma|indexlist_tma" -f -o gpurun_out/exp5_prof python tools/prof_kernels.py gemm indexlist > gpurun_out/exp5_prof.log 2>&1; echo "ncu rc=$?"
