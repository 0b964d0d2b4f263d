#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_r02_mgpu.sh "2 4 8"'    (rank counts to run on this box, each <= N)
# Multi-GPU evidence that survives the call (everything goes to gpurun_out/r02_mgpu_*):
#   1. tests/mgpu_check.py: the NVLink exchange (both launch forms) against the CPU simulation, sharded DOT / REDUCE_SUM vs oracle;
#   2. the reference's OWN driver with Base_B200 integrated on P ranks (MPI stand-in for the rendezvous):
#      Comm_HALO_EXCHANGE_FUSED, Base_Seq / Base_CUDA / Base_B200 in its cross-rank checksum report (Executor.cpp:1392-1467)
#      and its timing report -- Base_CUDA over the stand-in's shared-memory transport is the exchange incumbent;
#   3. this repo's exchange timings per launch form at 512^3 and 1024^3 per GPU, verified (tools/mgpu_halo.py).
RANKS=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mgpu_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02_mgpu_smi.txt 2>&1
for P in $RANKS; do
  TAG=r02_mgpu_n$P
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29611 \
      tests/mgpu_check.py > gpurun_out/${TAG}_check.log 2>&1; echo "mgpu_check P=$P rc=$?"
  grep -E "MGPU_CHECK|mgpu_check|mismatch" gpurun_out/${TAG}_check.log | head -5
  if [ -x oracle/_ref/raja-perf-with-b200-mpi.exe ]; then
    # Base_Seq in the report at 256^3 per rank (its host-side init + checksum is single-threaded per rank) ...
    timeout 400 python tools/mpirun_stub.py -n $P --gpu-per-rank -- oracle/_ref/raja-perf-with-b200-mpi.exe \
        -k Comm_HALO_EXCHANGE_FUSED -v Base_Seq Base_CUDA Base_B200 --checkrun 20 --size 16777216 \
        --outdir gpurun_out/${TAG}_ref_256 > gpurun_out/${TAG}_ref_256.log 2>&1; echo "reference driver 256^3, $P ranks rc=$?"
    grep -v "^$" gpurun_out/${TAG}_ref_256/RAJAPerf-checksum.txt | tail -12
    cat gpurun_out/${TAG}_ref_256/RAJAPerf-timing-Average.csv
    # ... and the GPU variants alone at 512^3 per rank, 100 reps: the same-size exchange incumbent
    timeout 400 python tools/mpirun_stub.py -n $P --gpu-per-rank -- oracle/_ref/raja-perf-with-b200-mpi.exe \
        -k Comm_HALO_EXCHANGE_FUSED -v Base_CUDA Base_B200 --checkrun 100 --size 134217728 \
        --outdir gpurun_out/${TAG}_ref_512 > gpurun_out/${TAG}_ref_512.log 2>&1; echo "reference driver 512^3, $P ranks rc=$?"
    grep -v "^$" gpurun_out/${TAG}_ref_512/RAJAPerf-checksum.txt | tail -10
    cat gpurun_out/${TAG}_ref_512/RAJAPerf-timing-Average.csv
  fi
  for G in 512 1024; do
    G=$G timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29612 \
        tools/mgpu_halo.py 2>gpurun_out/${TAG}_halo_$G.err | grep n_gpus | tee gpurun_out/${TAG}_halo_$G.json
  done
done
# the single-rank reference points on the same box
for G in 512 1024; do
  G=$G timeout 200 python tools/mgpu_halo.py 2>gpurun_out/r02_mgpu_n1_halo_$G.err | grep n_gpus | tee gpurun_out/r02_mgpu_n1_halo_$G.json
done
