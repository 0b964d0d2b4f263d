#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_d_mgpu_halo_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_d_mgpu_halo_n4.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_d_mgpu_halo_n2.json
timeout 300 python tools/mgpu_halo.py 2>/dev/null | grep n_gpus | tee gpurun_out/r01_d_mgpu_halo_n1.json
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r01_d_bench_n$N.json 2> gpurun_out/r01_d_bench_n$N.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r01_d_bench_n$N.json | grep -E "value|HALO|halo_exchange"
