#!/bin/bash
# Second part of the incumbent column (bounded box time): Apps, Comm, GEMM, SORT groups in --fast mode.
TAG=${TAG:-r01_h}
mkdir -p gpurun_out
timeout 175 python tools/incumbent_suite.py --fast --groups mass pa ltimes comm gemm sort --budget 105 --timeout 40 \
    --out gpurun_out/${TAG}_incumbent_suite2 > gpurun_out/${TAG}_incumbent_suite2.log 2>&1; echo "incumbent rc=$?"
cat gpurun_out/${TAG}_incumbent_suite2.log
