#!/bin/bash
# usage: gpurun --gpus 2 -- 'bash tools/gpu_r02_mgpu2b.sh'  -- the progressive one-launch exchange: parity on 1 and 2 ranks, timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_comm_gpu.py -m gpu -q -x > gpurun_out/r02_mgpu2b_pytest_comm.log 2>&1; echo "pytest comm rc=$?"; tail -2 gpurun_out/r02_mgpu2b_pytest_comm.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    tests/mgpu_check.py > gpurun_out/r02_mgpu2b_n2_check.log 2>&1; echo "mgpu_check P=2 rc=$?"
grep -E "MGPU_CHECK|mismatch" gpurun_out/r02_mgpu2b_n2_check.log | head -5
for G in 512 1024; do
  G=$G timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
      tools/mgpu_halo.py 2>gpurun_out/r02_mgpu2b_n2_halo_$G.err | grep n_gpus | tee gpurun_out/r02_mgpu2b_n2_halo_$G.json
done
G=512 timeout 200 python tools/mgpu_halo.py 2>gpurun_out/r02_mgpu2b_n1_halo_512.err | grep n_gpus | tee gpurun_out/r02_mgpu2b_n1_halo_512.json
