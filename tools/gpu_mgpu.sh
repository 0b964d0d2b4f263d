#!/bin/bash
# usage: tools/gpu_mgpu.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r01_d_bench_n$N.json 2> gpurun_out/r01_d_bench_n$N.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r01_d_bench_n$N.json | grep -E "value|HALO|halo_exchange"
