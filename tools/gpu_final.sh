#!/bin/bash
# Evidence at HEAD in one gpurun call (1 GPU): GPU parity tests, smoke, the bench line (with the cpu_baseline leg).
TAG=${TAG:-r01_i}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 150 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.json | head -40
timeout 40 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
