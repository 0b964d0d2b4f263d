#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_f_mgpu_halo_n2.json
timeout 300 python tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_f_mgpu_halo_n1.json
