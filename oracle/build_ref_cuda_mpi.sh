#!/bin/bash
# Authoring-container only: the UNMODIFIED reference WITH its CUDA variants (sm_100) AND its MPI code paths, linked against
# the MPI stand-in of oracle/mpi_stub/ -> oracle/_ref/raja-perf-cuda-mpi1.exe.  It adds the Base_CUDA / RAJA_CUDA variants of
# Comm_HALO_EXCHANGE, Comm_HALO_EXCHANGE_FUSED and Comm_HALO_SENDRECV to the incumbent column (one rank: every message is a
# self-send, a host memcpy in any MPI).  ~25 min on 8 cores; see build_ref_cuda.sh and build_ref_mpi.sh for the two halves.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
BUILD=${BUILD:-/tmp/rpb_refcudampi}
mkdir -p "$BUILD" "$HERE/_ref"
/usr/bin/gcc -O2 -fPIC -std=gnu11 -pthread -c "$HERE/mpi_stub/mpi_stub.c" -o "$BUILD/mpi_stub.o"
ar rcs "$BUILD/libmpistub.a" "$BUILD/mpi_stub.o"
cd "$BUILD"
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DENABLE_OPENMP=On -DENABLE_CUDA=On \
  -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc -DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++ \
  "-DCMAKE_CUDA_ARCHITECTURES=90-virtual;100-real" -DENABLE_TESTS=Off -DENABLE_MPI=On -DENABLE_FIND_MPI=Off \
  "-DBLT_MPI_INCLUDES=$HERE/mpi_stub" "-DBLT_MPI_LIBRARIES=$BUILD/libmpistub.a" /root/reference > cmake.log 2>&1
ninja raja-perf.exe > ninja.log 2>&1
cp bin/raja-perf.exe "$HERE/_ref/raja-perf-cuda-mpi1.exe"
echo "built $HERE/_ref/raja-perf-cuda-mpi1.exe"
