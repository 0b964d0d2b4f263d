#!/bin/bash
# Authoring-container only (the incumbent column): builds the UNMODIFIED reference WITH its CUDA
# variants (Base_CUDA / RAJA_CUDA and their cub tunings, CUB as vendored by the reference) for sm_100, out of tree
# from /root/reference with the reference's own CMake + Ninja, and drops the binary in
# oracle/_ref/raja-perf-cuda.exe (git-ignored, travels to the GPU box).  tools/incumbent_suite.py times it next to
# this repo's harness on the same B200.  nvcc cross-compiles without a GPU; ~25 min on 8 cores.
# "90-virtual;100-real": a plain "100" trips RAJA's architecture string compare (tpl/RAJA/CMakeLists.txt:116-119).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
BUILD=${BUILD:-/tmp/rpb_refcuda}
mkdir -p "$BUILD" "$HERE/_ref"
cd "$BUILD"
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DENABLE_OPENMP=On -DENABLE_CUDA=On \
  -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc -DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++ \
  "-DCMAKE_CUDA_ARCHITECTURES=90-virtual;100-real" -DENABLE_TESTS=Off /root/reference > cmake.log 2>&1
ninja raja-perf.exe > ninja.log 2>&1
cp bin/raja-perf.exe "$HERE/_ref/raja-perf-cuda.exe"
echo "built $HERE/_ref/raja-perf-cuda.exe"
