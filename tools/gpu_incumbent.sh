#!/bin/bash
# One gpurun call (1 GPU): the LTIMES fragment-mapping A/B, the harness-tunings parity test, and the incumbent column
# (reference Base_CUDA / RAJA_CUDA / Base_OpenMP next to Base_B200).  Outputs under gpurun_out/.
TAG=${TAG:-r01_h}
mkdir -p gpurun_out
timeout 100 python tools/time_quick.py ltimes_line > gpurun_out/${TAG}_ltimes_line.log 2>&1; echo "ltimes rc=$?"
grep -E "ltimes" gpurun_out/${TAG}_ltimes_line.log | tail -8
timeout 100 python -m pytest tests/test_suite_harness.py tests/test_apps_gpu.py -q -m gpu -k "tuning or ltimes" > gpurun_out/${TAG}_pytest_tunings.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest_tunings.log
timeout 270 python tools/incumbent_suite.py --budget 170 --timeout 60 --out gpurun_out/${TAG}_incumbent_suite > gpurun_out/${TAG}_incumbent_suite.log 2>&1; echo "incumbent rc=$?"
cat gpurun_out/${TAG}_incumbent_suite.log
