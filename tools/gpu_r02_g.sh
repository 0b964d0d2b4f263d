#!/bin/bash
# Round 2, call g (1 GPU): SORT with the 4-instruction match step, look-back batch 1 / 2 / 4, next to CUB on the same box.
TAG=${TAG:-r02_g}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_algorithm_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest_algo.log 2>&1; echo "pytest algorithm rc=$?"; tail -2 gpurun_out/${TAG}_pytest_algo.log
timeout 300 python tools/time_r02.py sort --out gpurun_out/${TAG}_time.json > gpurun_out/${TAG}_time.log 2>&1; echo "time_r02 rc=$?"
cat gpurun_out/${TAG}_time.log
[ -x tools/bin/incumbent ] && (timeout 120 tools/bin/incumbent > gpurun_out/${TAG}_cub.jsonl 2>&1; head -2 gpurun_out/${TAG}_cub.jsonl)
