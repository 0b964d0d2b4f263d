#!/bin/bash
# The first gpurun call of the next round (1 GPU, ~8 min of box time): everything round 1 left unmeasured.
#   1. the two opt-in tunings written after the GPU budget was spent: parity + A/B (SCAN line-major stores, SORT lane-private
#      histogram) -- if they win, make them the defaults (csrc/scan_tma.cu ST_LAUNCH, csrc/sort.cu sort_impl);
#   2. ncu --set full of the kernels whose defaults changed after the last capture (MASS3DPA / LTIMES line-major accesses):
#      DRAM bytes per launch, L2 request counts (lts__t_requests) -- the evidence behind profiles/r01_pa_variants.md;
#   3. the incumbent column of the three MPI-only kernels (reference Base_CUDA over the MPI stand-in, one rank);
#   4. the full evidence at HEAD: pytest -m gpu, bench.py (now with the per-kernel cpu_openmp column), smoke.
TAG=${TAG:-r02_a}
mkdir -p gpurun_out
timeout 120 python tools/time_quick.py scan_line sort_hist > gpurun_out/${TAG}_optin.log 2>&1; echo "opt-in A/B rc=$?"
grep -E "parity|scan |sort keys" gpurun_out/${TAG}_optin.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"mass3dpa_kernel|ltimes_dmma_kernel|scan_tma_kernel" -c 4 -f \
    -o gpurun_out/${TAG}_prof python tools/prof_kernels.py pa ltimes scan > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
timeout 90 python tools/incumbent_suite.py --fast --groups exchange --timeout 40 --out gpurun_out/${TAG}_incumbent_exchange \
    > gpurun_out/${TAG}_incumbent_exchange.log 2>&1; echo "incumbent exchange rc=$?"; tail -6 gpurun_out/${TAG}_incumbent_exchange.log
# 5. the reference's OWN driver with the Base_B200 variant integrated (rajaperf_b200/integration/): its checksum report
#    compares Base_B200 with Base_Seq / Base_CUDA itself, at the suite's default sizes
if [ -x oracle/_ref/raja-perf-with-b200.exe ]; then
  timeout 120 oracle/_ref/raja-perf-with-b200.exe -k Stream Algorithm_REDUCE_SUM Algorithm_SCAN Algorithm_SORT Algorithm_SORTPAIRS \
      Apps_MASS3DPA Apps_DIFFUSION3DPA Apps_CONVECTION3DPA Apps_LTIMES Comm_HALO_PACKING_FUSED \
      -v Base_Seq Base_CUDA RAJA_CUDA Base_B200 --checkrun 5 --outdir gpurun_out/${TAG}_ref_with_b200 \
      > gpurun_out/${TAG}_ref_with_b200.log 2>&1; echo "reference driver with Base_B200 rc=$?"
  grep -E "^(Stream|Algorithm|Apps|Comm)_|Base_B200|Base_Seq" gpurun_out/${TAG}_ref_with_b200/RAJAPerf-checksum.txt | head -60
fi
TAG=$TAG bash tools/gpu_final.sh
