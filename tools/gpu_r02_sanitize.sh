#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled synchronisation: racecheck (shared-memory hazards),
# synccheck (barrier misuse), memcheck (out-of-bounds / misaligned).  Logs -> gpurun_out/r02_sanitize_<tool>_<target>.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  for target in reduce scan indexlist pa sort halo; do
    log=gpurun_out/r02_sanitize_${tool}_${target}.log
    timeout ${SAN_TIMEOUT:-240} $CS --tool $tool --print-limit 20 python tools/sanitize_targets.py $target > $log 2>&1
    rc=$?
    echo "$tool $target rc=$rc | $(grep -E 'SANITIZE_TARGETS|ERROR SUMMARY|RACECHECK SUMMARY' $log | tr '\n' ' ')"
  done
done
