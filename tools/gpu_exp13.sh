#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_suite_harness.py -m gpu -x -q -k "sendrecv or exchange" 2>&1 | tail -8
timeout 600 python tools/time_quick.py pa > gpurun_out/exp13_time.log 2>&1; cat gpurun_out/exp13_time.log
cd rajaperf_b200/suite && ./raja-perf-b200.exe -k HALO_SENDRECV HALO_EXCHANGE_FUSED --size 134217728 --checkrun 50 --graph --outdir /tmp/halo_out > /dev/null 2>&1; cat /tmp/halo_out/RAJAPerf-bandwidth.csv
