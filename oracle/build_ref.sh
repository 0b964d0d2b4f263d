#!/bin/bash
# Authoring-container only: builds the UNMODIFIED reference (CPU-only: Base_Seq / Base_OpenMP / RAJA_OpenMP
# variants) out of tree from /root/reference with the reference's own CMake + Ninja, and drops the binary in
# oracle/_ref/raja-perf.exe (git-ignored).  tests/golden/make_golden.py runs it to mint the known-answer
# checksums in tests/golden/ref_checksums.json; when present it is also the "reference" CPU leg of bench.py.
# It is NOT part of __graft_entry__.build(): the build needs cmake/BLT and generated headers (~2 min on 8 cores).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
BUILD=${BUILD:-/tmp/rpb_refbuild}
mkdir -p "$BUILD" "$HERE/_ref"
cd "$BUILD"
# the image's default CC/CXX wrappers have no libgomp.spec: use the system compilers
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DENABLE_OPENMP=On -DENABLE_CUDA=Off \
  -DENABLE_TESTS=Off /root/reference > cmake.log 2>&1
ninja raja-perf.exe > ninja.log 2>&1
cp bin/raja-perf.exe "$HERE/_ref/raja-perf.exe"
echo "built $HERE/_ref/raja-perf.exe"
