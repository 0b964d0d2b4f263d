#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widened_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/time_quick.py indexlist gemm > gpurun_out/exp6_time.log 2>&1; cat gpurun_out/exp6_time.log
