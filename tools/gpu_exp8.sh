#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_algorithm_gpu.py -m gpu -x -q -k scan 2>&1 | tail -3
: > gpurun_out/exp8_time.log
for ds in 1 2 4 8 16 32; do
  RPB200_SCAN_DSTRIDE=$ds RPB200_IL_DSTRIDE=$((ds*4)) timeout 120 python tools/time_scan_il.py >> gpurun_out/exp8_time.log 2>&1
done
cat gpurun_out/exp8_time.log
