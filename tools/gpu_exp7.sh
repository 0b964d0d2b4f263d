#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_widened_gpu.py -m gpu -x -q -k indexlist 2>&1 | tail -3
: > gpurun_out/exp7_time.log
for ds in 1 4 16; do for lbw in 1 2 4; do for bo in 0 100; do
  echo "== dstride=$ds lbw=$lbw backoff=$bo" >> gpurun_out/exp7_time.log
  RPB200_IL_DSTRIDE=$ds RPB200_IL_LBW=$lbw RPB200_IL_BACKOFF=$bo timeout 120 python tools/time_quick.py indexlist 2>&1 | grep "cps=4" >> gpurun_out/exp7_time.log
done; done; done
cat gpurun_out/exp7_time.log
