#!/bin/bash
# One gpurun call: MASS3DPA line-major A/B + parity, the Apps / harness GPU tests, a bench line without the CPU leg.
TAG=${TAG:-r01_j}
mkdir -p gpurun_out
timeout 60 python tools/time_quick.py mass_line > gpurun_out/${TAG}_mass_line.log 2>&1; echo "mass rc=$?"
grep -E "mass3dpa" gpurun_out/${TAG}_mass_line.log | tail -10
timeout 90 python -m pytest tests/test_apps_gpu.py tests/test_suite_harness.py -x -q -m gpu  > gpurun_out/${TAG}_pytest_apps.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest_apps.log
timeout 100 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.json | grep -E "value|MASS|LTIMES|SCAN"
timeout 40 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
