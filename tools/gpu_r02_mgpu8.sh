#!/bin/bash
# usage: gpurun --gpus 8 -- 'bash tools/gpu_r02_mgpu8.sh'
# The 4- and 8-rank evidence in one call (see tools/gpu_r02_mgpu.sh for what each step is): mgpu_check on 8 ranks; the reference's
# own driver with Base_B200 integrated on 4 ranks (256^3: Base_Seq / Base_CUDA / Base_B200) and on 8 ranks (256^3 and 512^3);
# this repo's exchange timings (two-launch forms, pack / unpack split, NVLink counters) at 512^3 on 1 / 4 / 8 ranks and at
# 1024^3 on 8.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mgpu8_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02_mgpu8_smi.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    tests/mgpu_check.py > gpurun_out/r02_mgpu_n8_check.log 2>&1; echo "mgpu_check P=8 rc=$?"
grep -E "MGPU_CHECK|mismatch" gpurun_out/r02_mgpu_n8_check.log | head -5
for P in 4 8; do
  TAG=r02_mgpu_n$P
  timeout 400 python tools/mpirun_stub.py -n $P --gpu-per-rank -- oracle/_ref/raja-perf-with-b200-mpi.exe \
      -k Comm_HALO_EXCHANGE_FUSED -v Base_Seq Base_CUDA Base_B200 --checkrun 20 --size 16777216 \
      --outdir gpurun_out/${TAG}_ref_256 > gpurun_out/${TAG}_ref_256.log 2>&1; echo "reference driver 256^3, $P ranks rc=$?"
  grep -v "^$" gpurun_out/${TAG}_ref_256/RAJAPerf-checksum.txt | tail -5
  tail -2 gpurun_out/${TAG}_ref_256/RAJAPerf-timing-Average.csv
done
timeout 400 python tools/mpirun_stub.py -n 8 --gpu-per-rank -- oracle/_ref/raja-perf-with-b200-mpi.exe \
    -k Comm_HALO_EXCHANGE_FUSED -v Base_CUDA Base_B200 --checkrun 100 --size 134217728 \
    --outdir gpurun_out/r02_mgpu_n8_ref_512 > gpurun_out/r02_mgpu_n8_ref_512.log 2>&1; echo "reference driver 512^3, 8 ranks rc=$?"
grep -v "^$" gpurun_out/r02_mgpu_n8_ref_512/RAJAPerf-checksum.txt | tail -4
tail -2 gpurun_out/r02_mgpu_n8_ref_512/RAJAPerf-timing-Average.csv
for P in 8 4; do
  FORMS=two G=512 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29612 \
      tools/mgpu_halo.py 2>gpurun_out/r02_mgpu_n${P}_halo_512.err | grep n_gpus | tee gpurun_out/r02_mgpu_n${P}_halo_512.json
done
FORMS=two G=1024 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
    tools/mgpu_halo.py 2>gpurun_out/r02_mgpu_n8_halo_1024.err | grep n_gpus | tee gpurun_out/r02_mgpu_n8_halo_1024.json
FORMS=two G=512 timeout 120 python tools/mgpu_halo.py 2>gpurun_out/r02_mgpu8_n1_halo_512.err | grep n_gpus | tee gpurun_out/r02_mgpu8_n1_halo_512.json
