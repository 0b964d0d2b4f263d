#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_suite_harness.py tests/test_apps_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/time_quick.py pa > gpurun_out/exp12_time.log 2>&1; cat gpurun_out/exp12_time.log
cd rajaperf_b200/suite && ./raja-perf-b200.exe -k HALO_PACKING HALO_PACKING_FUSED HALO_EXCHANGE HALO_EXCHANGE_FUSED --size 134217728 --checkrun 50 --graph --outdir /tmp/halo_out > /dev/null 2>&1; cat /tmp/halo_out/RAJAPerf-bandwidth.csv; cat /tmp/halo_out/RAJAPerf-timing-Average.csv
./raja-perf-b200.exe -k HALO_PACKING HALO_PACKING_FUSED HALO_EXCHANGE HALO_EXCHANGE_FUSED --size 134217728 --checkrun 50 --outdir /tmp/halo_out2 > /dev/null 2>&1; cat /tmp/halo_out2/RAJAPerf-timing-Average.csv
