#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_f_mgpu_halo_n8.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_f_mgpu_halo_n4.json
