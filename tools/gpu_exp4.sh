#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/time_quick.py indexlist gemm > gpurun_out/exp4_time.log 2>&1; cat gpurun_out/exp4_time.log
