#!/bin/bash
# Authoring-container only: the reference suite WITH the Base_B200 variant integrated (apply_base_b200.py: a patched COPY of
# /root/reference under /tmp, nothing is stored in this repo), built with the reference's own CMake for sm_100 and linked
# against rajaperf_b200/lib/librpb200.so -> oracle/_ref/raja-perf-with-b200.exe.  On a B200:
#   oracle/_ref/raja-perf-with-b200.exe -k Stream Algorithm_SCAN Apps_MASS3DPA -v Base_Seq Base_CUDA Base_B200 --checkrun 5
# prints the reference's OWN checksum report (Base_B200 against Base_Seq, test/test-raja-perf-suite.cpp:124-167) and its own
# timing report with the three variants side by side.  ~25 min on 8 cores.  The run path of librpb200.so is absolute
# ($ROOT/rajaperf_b200/lib: /root/repo/... exists on the GPU box too, as a symlink to the snapshot).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
SRC=${SRC:-/tmp/ref_b200}
BUILD=${BUILD:-/tmp/ref_b200_build}
[ -d "$SRC/src" ] || python "$HERE/apply_base_b200.py" --dst "$SRC"
mkdir -p "$BUILD" "$ROOT/oracle/_ref"
# MPI=1: also the MPI code paths (Comm_HALO_EXCHANGE_FUSED with its Base_B200 stub) against the MPI stand-in of
# oracle/mpi_stub -> raja-perf-with-b200-mpi.exe; P ranks = P processes launched like tests/golden/make_golden.py:mpirun,
# one GPU each (CUDA_VISIBLE_DEVICES per rank)
MPI_FLAGS=()
OUT=raja-perf-with-b200.exe
if [ "${MPI:-0}" = "1" ]; then
  /usr/bin/gcc -O2 -fPIC -std=gnu11 -pthread -c "$ROOT/oracle/mpi_stub/mpi_stub.c" -o "$BUILD/mpi_stub.o"
  ar rcs "$BUILD/libmpistub.a" "$BUILD/mpi_stub.o"
  MPI_FLAGS=(-DENABLE_MPI=On -DENABLE_FIND_MPI=Off "-DBLT_MPI_INCLUDES=$ROOT/oracle/mpi_stub" "-DBLT_MPI_LIBRARIES=$BUILD/libmpistub.a")
  OUT=raja-perf-with-b200-mpi.exe
fi
# TESTS=1: also the reference's own gtest (test/test-raja-perf-suite.cpp: every kernel, every variant that ran, checksum within
# 1e-7 of the first one) -> oracle/_ref/test-raja-perf-suite-with-b200.exe; turning it on in an existing build tree only adds
# gtest and the test target (6 build steps)
TESTS_FLAG=-DENABLE_TESTS=Off
[ "${TESTS:-0}" = "1" ] && TESTS_FLAG=-DENABLE_TESTS=On
cd "$BUILD"
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DENABLE_OPENMP=On -DENABLE_CUDA=On \
  -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc -DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++ \
  "-DCMAKE_CUDA_ARCHITECTURES=90-virtual;100-real" $TESTS_FLAG \
  "-DCMAKE_CXX_FLAGS=-I$ROOT/include" "-DCMAKE_CUDA_FLAGS=-I$ROOT/include" \
  "-DCMAKE_CXX_STANDARD_LIBRARIES=-L$ROOT/rajaperf_b200/lib -lrpb200 -Wl,-rpath,$ROOT/rajaperf_b200/lib" \
  "${MPI_FLAGS[@]}" "$SRC" > cmake.log 2>&1
ninja raja-perf.exe > ninja.log 2>&1
cp bin/raja-perf.exe "$ROOT/oracle/_ref/$OUT"
echo "built $ROOT/oracle/_ref/$OUT"
if [ "${TESTS:-0}" = "1" ] && [ "${MPI:-0}" != "1" ]; then
  ninja test-raja-perf-suite.exe >> ninja.log 2>&1
  cp test/test-raja-perf-suite.exe "$ROOT/oracle/_ref/test-raja-perf-suite-with-b200.exe"
  echo "built $ROOT/oracle/_ref/test-raja-perf-suite-with-b200.exe"
fi
