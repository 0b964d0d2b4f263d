#!/bin/bash
# Authoring-container only: builds the UNMODIFIED reference CPU-only WITH its MPI code paths, linked against the
# MPI stand-in of oracle/mpi_stub/ (this image has no MPI; one rank in process, or P processes over shared memory), out of tree from /root/reference with the
# reference's own CMake + Ninja, and drops the binary in oracle/_ref/raja-perf-mpi1.exe (git-ignored).  With MPI enabled
# the suite compiles Comm_HALO_EXCHANGE, Comm_HALO_EXCHANGE_FUSED and Comm_HALO_SENDRECV in (RAJAPerfSuite.hpp:177-181);
# tests/golden/make_golden.py --mpi mints the exchange goldens (tests/golden/ref_checksums_mpi1.json) from it, on 1 rank
# and on 2 / 4 / 6 / 8 ranks.  ~2 min on 8 cores.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
BUILD=${BUILD:-/tmp/rpb_refmpi}
mkdir -p "$BUILD" "$HERE/_ref"
/usr/bin/gcc -O2 -fPIC -std=gnu11 -pthread -c "$HERE/mpi_stub/mpi_stub.c" -o "$BUILD/mpi_stub.o"
ar rcs "$BUILD/libmpistub.a" "$BUILD/mpi_stub.o"
cd "$BUILD"
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DENABLE_OPENMP=On -DENABLE_CUDA=Off \
  -DENABLE_TESTS=Off -DENABLE_MPI=On -DENABLE_FIND_MPI=Off "-DBLT_MPI_INCLUDES=$HERE/mpi_stub" \
  "-DBLT_MPI_LIBRARIES=$BUILD/libmpistub.a" /root/reference > cmake.log 2>&1
ninja raja-perf.exe > ninja.log 2>&1
cp bin/raja-perf.exe "$HERE/_ref/raja-perf-mpi1.exe"
echo "built $HERE/_ref/raja-perf-mpi1.exe"
