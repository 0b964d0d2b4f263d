#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_apps_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/time_quick.py pa > gpurun_out/exp17_time.log 2>&1; cat gpurun_out/exp17_time.log
