#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_round2_mgpu.sh N'   (N = 2, 4 or 8)
# The reference's OWN driver with Base_B200 integrated, on N ranks (one process and one GPU per rank, MPI stand-in transport
# for the rendezvous): Comm_HALO_EXCHANGE_FUSED, Base_B200 next to Base_Seq and Base_CUDA in the reference's checksum and
# timing reports.  Base_CUDA's messages travel through the stand-in's shared-memory log (host memcpy): a proxy for an MPI
# without CUDA-aware transport, not a tuned MPI.  Then this repo's own N-GPU measurement (tools/mgpu_halo.py).
N=${1:-2}
TAG=${TAG:-r02_mgpu_n$N}
mkdir -p gpurun_out
if [ -x oracle/_ref/raja-perf-with-b200-mpi.exe ]; then
  timeout 300 python tools/mpirun_stub.py -n $N --gpu-per-rank -- oracle/_ref/raja-perf-with-b200-mpi.exe \
      -k Comm_HALO_EXCHANGE_FUSED -v Base_Seq Base_CUDA Base_B200 --checkrun 20 --size 16777216 \
      --outdir gpurun_out/${TAG}_ref_with_b200 > gpurun_out/${TAG}_ref_with_b200.log 2>&1; echo "reference driver, $N ranks rc=$?"
  grep -v "^$" gpurun_out/${TAG}_ref_with_b200/RAJAPerf-checksum.txt | tail -8
  cat gpurun_out/${TAG}_ref_with_b200/RAJAPerf-timing-Average.csv
fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/${TAG}_mgpu_halo.json
