#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_widened_gpu.py -m gpu -x -q -k indexlist 2>&1 | tail -3
for nlb in 3 1; do RPB200_IL_NLB=$nlb timeout 120 python tools/time_scan_il.py 2>&1 | grep indexlist; done
