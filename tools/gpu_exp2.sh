#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_algorithm_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/time_quick.py sort > gpurun_out/exp2_time.log 2>&1; tail -4 gpurun_out/exp2_time.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__t_requests.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum -k regex:halo_kernel --csv --log-file gpurun_out/exp2_halo_l2.csv python tools/prof_halo_l2.py > gpurun_out/exp2_halo.log 2>&1; echo "ncu rc=$?"
G=512 timeout 300 python tools/halo_faces.py > gpurun_out/exp2_faces.log 2>&1; cat gpurun_out/exp2_faces.log
