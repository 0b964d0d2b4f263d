#!/bin/bash
# tools/gpu_call.sh -- ONE parameterised script for every evidence call on a B200 box (replaces the per-call gpu_*.sh scripts
# of rounds 1 and 2; the calls those made are listed at the bottom as step lists).
#
#   gpurun [--gpus N] --timeout S -- 'TAG=r03_a bash tools/gpu_call.sh STEP [STEP ...]'
#   DRY=1 bash tools/gpu_call.sh STEP ...          # print the commands instead of running them (works without a GPU)
#
# Every step writes gpurun_out/${TAG}_<what>.{log,json,csv} and echoes "<step> rc=<rc>" plus a short tail, so the call's
# stdout tail is a readable verdict.  A step is NAME or NAME:ARG[:ARG...]; ARGs with several words use '+' for the blank
# (tests:tests/test_comm_gpu.py, time:halo+sort, mgpu_ref:8:134217728:100:Base_CUDA+Base_B200).
#
#   smi                      GPU name / clocks / topology
#   tests[:PATHS]            pytest -m gpu (-x), whole suite or PATHS
#   smoke                    __graft_entry__.smoke()
#   bench[:FLAGS]            python bench.py FLAGS  (default: full line with the CPU leg)
#   bench_ref                python bench.py --impl reference --steps 3 --warmup 1
#   bench_n:P[:FLAGS]        torchrun -n P bench.py --gpus P FLAGS (default --steps 10 --warmup 3 --no-cpu)
#   launches                 ncu launch list (gpu__time_duration.sum) of `bench.py --steps 2 --warmup 3 --no-cpu --no-extras`
#   ncu:SECTIONS             ncu --set full of tools/prof_kernels.py SECTIONS (+ raw csv + launch manifest, which tools/ncu_traffic.py
#                            turns into profiles/rNN_ncu_traffic.json back in the authoring container)
#   ncu_halo                 ncu --set full of the halo launches at HEAD (tools/prof_halo_r02.py) + the DRAM/L2 table
#   time:SECTIONS            tools/time_r02.py SECTIONS   (halo halo1024 sort reduce)
#   time_quick:SECTIONS      tools/time_quick.py SECTIONS (scan pa mass_line ltimes ltimes_line halo sort indexlist gemm)
#   cub                      tools/bin/incumbent: cub::DeviceRadixSort / DeviceScan / DeviceReduce / cudaMemcpy on this box
#   sanitize[:TARGETS]       compute-sanitizer memcheck + synccheck + racecheck over tools/sanitize_targets.py TARGETS
#                            (default: reduce scan indexlist pa sort halo)
#   adjudicate:PHASES[:KERNELS]   tools/ref_adjudicate.py (the reference's own driver with Base_B200 integrated)
#   gtest                    the reference's own gtest built with Base_B200 integrated
#   harness                  the C++ suite harness over all kernels at the BASELINE sizes (--graph), one bandwidth csv
#   incumbent_suite[:GROUPS] tools/incumbent_suite.py --fast [--groups GROUPS]
#   mgpu_tests               pytest tests/test_multigpu.py -m gpu (needs >= 2 GPUs)
#   mgpu_check:P             torchrun -n P tests/mgpu_check.py (exchange vs CPU simulation, sharded DOT / REDUCE_SUM vs oracle)
#   mgpu_ref:P:SIZE:REPS:VARIANTS   the reference driver (MPI stand-in) on P ranks, one GPU each: Comm_HALO_EXCHANGE_FUSED
#   mgpu_halo:P:G[:FORMS]    tools/mgpu_halo.py on P ranks at G^3 cells per GPU (FORMS=two: two-launch forms only)
TAG=${TAG:-r03_a}
O=gpurun_out
PY=python
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p $O

run() {            # run NAME LIMIT_S OUTFILE CMD...   (stdout+stderr -> OUTFILE)
  local name=$1 limit=$2 out=$3; shift 3
  if [ -n "$DRY" ]; then echo "[$name] timeout $limit $* > $out 2>&1"; return 0; fi
  timeout $limit "$@" > $out 2>&1; local rc=$?
  echo "$name rc=$rc"; return $rc
}
run2() {           # run2 NAME LIMIT_S OUT ERR CMD...  (stdout -> OUT, stderr -> ERR: for JSON lines)
  local name=$1 limit=$2 out=$3 err=$4; shift 4
  if [ -n "$DRY" ]; then echo "[$name] timeout $limit $* > $out 2> $err"; return 0; fi
  timeout $limit "$@" > $out 2> $err; local rc=$?
  echo "$name rc=$rc"; return $rc
}
words() { echo "${1//+/ }"; }
show() { [ -n "$DRY" ] || "$@"; }

halo_table() {     # the DRAM / L2 columns of an `ncu --page raw --csv` file, one row per launch
  show $PY - "$1" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
ix = [h.index(w) for w in want if w in h]
print([h[i] for i in ix]); print([rows[1][i] for i in ix])
for r in rows[2:]:
    print([r[i][:40] for i in ix])
PY
}

step() {
  local IFS=:; set -- $1; unset IFS
  local s=$1 a=$2 b=$3 c=$4 d=$5
  case $s in
    smi)
      show bash -c "nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1; nvidia-smi -L >> $O/${TAG}_smi.txt; nvidia-smi topo -m >> $O/${TAG}_smi.txt 2>&1; free -g | head -2; nproc" ;;
    tests)
      run tests 900 $O/${TAG}_pytest_gpu.log $PY -m pytest $(words "${a:-tests}") -m gpu -x -q; show tail -4 $O/${TAG}_pytest_gpu.log ;;
    smoke)
      run smoke 300 $O/${TAG}_smoke.log $PY __graft_entry__.py smoke; show tail -1 $O/${TAG}_smoke.log ;;
    bench)
      run2 bench 600 $O/${TAG}_bench.json $O/${TAG}_bench.err $PY bench.py $(words "$a"); show tail -3 $O/${TAG}_bench.err
      show bash -c "$PY tools/show_bench.py $O/${TAG}_bench.json | head -60" ;;
    bench_ref)
      run bench_ref 400 $O/${TAG}_bench_ref.json $PY bench.py --impl reference --steps 3 --warmup 1 ;;
    bench_n)
      run2 "bench_n$a" 900 $O/${TAG}_bench_n$a.json $O/${TAG}_bench_n$a.err $TORCHRUN --nproc-per-node $a --master-port 29611 \
          bench.py --gpus $a $(words "${b:---steps+10+--warmup+3+--no-cpu}")
      show bash -c "$PY tools/show_bench.py $O/${TAG}_bench_n$a.json | grep -E 'value|HALO|halo_exchange|sharded|verified'" ;;
    launches)
      run launches 600 $O/${TAG}_bench_under_ncu.log ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file $O/${TAG}_launches_bench.csv $PY bench.py --steps 2 --warmup 3 --no-cpu --no-extras ;;
    ncu)
      run ncu 900 $O/${TAG}_prof.log ncu --set full --clock-control none --import-source on \
          -k "regex:stream_ew|reduce_kernel|scan_tma|3dpa_kernel|ltimes|halo_|gemm_dmma|indexlist_tma|sort_" -c 24 -f -o $O/${TAG}_prof \
          $PY tools/prof_kernels.py $(words "${a:-stream scan indexlist pa ltimes sort}") --manifest $O/${TAG}_prof_manifest.json
      show bash -c "ncu -i $O/${TAG}_prof.ncu-rep --page raw --csv > $O/${TAG}_prof_raw.csv 2>/dev/null"
      # back in the authoring container: python tools/ncu_traffic.py gpurun_out/${TAG}_prof_raw.csv gpurun_out/${TAG}_prof_manifest.json
      ;;
    ncu_halo)
      run ncu_halo 300 $O/${TAG}_halo_ncu.log ncu --set full --cache-control none --clock-control none --import-source on \
          -k 'regex:halo_items_kernel|halo_kernel' -c 17 -f -o $O/${TAG}_halo $PY tools/prof_halo_r02.py
      show bash -c "ncu -i $O/${TAG}_halo.ncu-rep --page raw --csv > $O/${TAG}_halo_raw.csv 2>/dev/null"
      halo_table $O/${TAG}_halo_raw.csv ;;
    time)
      local tag2=${a//+/_}
      run "time $a" 400 $O/${TAG}_time_$tag2.log $PY tools/time_r02.py $(words "$a") --out $O/${TAG}_time_$tag2.json; show cat $O/${TAG}_time_$tag2.log ;;
    time_quick)
      run "time_quick $a" 300 $O/${TAG}_quick_${a//+/_}.log $PY tools/time_quick.py $(words "$a"); show tail -30 $O/${TAG}_quick_${a//+/_}.log ;;
    cub)
      if [ -x tools/bin/incumbent ] || [ -n "$DRY" ]; then run cub 120 $O/${TAG}_cub.jsonl tools/bin/incumbent; show cat $O/${TAG}_cub.jsonl; fi ;;
    sanitize)
      for tool in memcheck synccheck racecheck; do
        for target in $(words "${a:-reduce scan indexlist pa sort halo}"); do
          local log=$O/${TAG}_sanitize_${tool}_${target}.log
          run "$tool $target" ${SAN_TIMEOUT:-240} $log /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 \
              $PY tools/sanitize_targets.py $target
          show bash -c "grep -E 'SANITIZE_TARGETS|ERROR SUMMARY|RACECHECK SUMMARY' $log | tr '\n' ' '; echo"
        done
      done ;;
    adjudicate)
      local k=""; [ -n "$b" ] && k="--kernels $(words "$b")"
      run adjudicate 1500 $O/${TAG}_adjudicate.log $PY tools/ref_adjudicate.py --phases $(words "${a:-default checksum timing}") $k \
          --out $O/${TAG}_adjudicate; show tail -40 $O/${TAG}_adjudicate.log ;;
    gtest)
      if [ -x oracle/_ref/test-raja-perf-suite-with-b200.exe ] || [ -n "$DRY" ]; then
        if [ -n "$DRY" ]; then echo "[gtest] (cd $O && timeout 600 ../oracle/_ref/test-raja-perf-suite-with-b200.exe > ${TAG}_ref_gtest.log 2>&1)"
        else (cd $O && timeout 600 ../oracle/_ref/test-raja-perf-suite-with-b200.exe > ${TAG}_ref_gtest.log 2>&1; echo "reference gtest rc=$?")
          grep -E "^\[|Base_B200" $O/${TAG}_ref_gtest.log | grep -E "^\[|B200" | tail -25; fi
      fi ;;
    harness)
      local H=rajaperf_b200/suite/raja-perf-b200.exe D=$O/${TAG}_harness
      if [ -n "$DRY" ]; then echo "[harness] $H -k <group> --size <BASELINE size> --checkrun R --graph --outdir $D/<group>  (8 groups)"; return; fi
      rm -rf $D; mkdir -p $D
      ( timeout 150 $H -k Stream MEMCPY MEMSET --size 268435456 --checkrun 20 --graph --outdir $D/stream
        timeout 150 $H -k REDUCE_SUM SCAN INDEXLIST INDEXLIST_3LOOP --size 134217728 --checkrun 20 --graph --outdir $D/algo
        timeout 150 $H -k SORT SORTPAIRS --size 134217728 --checkrun 2 --graph --outdir $D/sort
        timeout 150 $H -k MASS3DPA --size 500000000 --checkrun 10 --graph --outdir $D/mass
        timeout 150 $H -k DIFFUSION3DPA CONVECTION3DPA --size 256000000 --checkrun 10 --graph --outdir $D/pa
        timeout 150 $H -k LTIMES --size 1024000000 --checkrun 10 --graph --outdir $D/ltimes
        timeout 150 $H -k Polybench_GEMM --size 16777216 --checkrun 5 --graph --outdir $D/gemm
        timeout 150 $H -k HALO_PACKING HALO_PACKING_FUSED HALO_SENDRECV HALO_EXCHANGE HALO_EXCHANGE_FUSED --size 134217728 \
            --checkrun 50 --graph --outdir $D/comm ) > $O/${TAG}_harness.log 2>&1
      ( head -1 $D/stream/RAJAPerf-bandwidth.csv; for g in stream algo sort mass pa ltimes gemm comm; do tail -n +2 $D/$g/RAJAPerf-bandwidth.csv; done ) \
          > $O/${TAG}_harness_bandwidth.csv
      cut -d, -f1,3,4,8,9,12 $O/${TAG}_harness_bandwidth.csv ;;
    incumbent_suite)
      local g=""; [ -n "$a" ] && g="--groups $(words "$a")"
      run incumbent_suite 300 $O/${TAG}_incumbent_suite.log $PY tools/incumbent_suite.py --fast $g --budget 200 --timeout 60 \
          --out $O/${TAG}_incumbent_suite; show cat $O/${TAG}_incumbent_suite.log ;;
    mgpu_tests)
      run mgpu_tests 600 $O/${TAG}_pytest_multigpu.log $PY -m pytest tests/test_multigpu.py -m gpu -x -q; show tail -3 $O/${TAG}_pytest_multigpu.log ;;
    mgpu_check)
      run "mgpu_check P=$a" 300 $O/${TAG}_n${a}_check.log $TORCHRUN --nproc-per-node $a --master-port 29611 tests/mgpu_check.py
      show bash -c "grep -E 'MGPU_CHECK|mgpu_check|mismatch' $O/${TAG}_n${a}_check.log | head -5" ;;
    mgpu_ref)
      if [ -x oracle/_ref/raja-perf-with-b200-mpi.exe ] || [ -n "$DRY" ]; then
        local D=$O/${TAG}_n${a}_ref_$b
        run "reference driver --size $b, $a ranks" 400 $D.log $PY tools/mpirun_stub.py -n $a --gpu-per-rank -- \
            oracle/_ref/raja-perf-with-b200-mpi.exe -k Comm_HALO_EXCHANGE_FUSED -v $(words "${d:-Base_Seq+Base_CUDA+Base_B200}") \
            --checkrun ${c:-20} --size $b --outdir $D
        show bash -c "grep -v '^$' $D/RAJAPerf-checksum.txt | tail -8; cat $D/RAJAPerf-timing-Average.csv"
      fi ;;
    mgpu_halo)
      local out=$O/${TAG}_n${a}_halo_$b
      if [ -n "$DRY" ]; then echo "[mgpu_halo] FORMS=$c G=$b $TORCHRUN --nproc-per-node $a tools/mgpu_halo.py > $out.json"; return; fi
      if [ "$a" = 1 ]; then FORMS=$c G=$b timeout 200 $PY tools/mgpu_halo.py 2> $out.err | grep n_gpus | tee $out.json
      else FORMS=$c G=$b timeout 300 $TORCHRUN --nproc-per-node $a --master-port 29612 tools/mgpu_halo.py 2> $out.err | grep n_gpus | tee $out.json; fi ;;
    *) echo "unknown step: $s"; return 2 ;;
  esac
}

for st in "$@"; do step "$st"; done

# The calls of rounds 1 and 2 as step lists (TAG in brackets):
#   [r01_i]  smi tests bench smoke                                              (gpu_final.sh)
#   [r01_e]  smi tests smoke bench bench_ref launches ncu:scan+indexlist+gemm+pa harness            (gpu_round.sh)
#   [r01_h]  time_quick:ltimes_line tests:tests/test_suite_harness.py+tests/test_apps_gpu.py incumbent_suite    (gpu_incumbent*.sh)
#   [r01_j]  time_quick:mass_line tests:tests/test_apps_gpu.py+tests/test_suite_harness.py bench:--no-cpu smoke  (gpu_mass.sh)
#   [r01_d]  mgpu_tests bench_n:N ; mgpu_halo:N:512                             (gpu_mgpu*.sh, --gpus N)
#   [r02_a]  smi time_quick:scan_line+sort_hist adjudicate incumbent_suite:exchange
#   [r02_b]  smi tests time:halo+sort+reduce time:halo1024 ncu_halo gtest adjudicate:timing:Stream_DOT+Algorithm_REDUCE_SUM+...
#   [r02_c/d] tests time:halo+sort cub time:halo1024 ncu_halo bench adjudicate:timing:Algorithm_SORT+Algorithm_SORTPAIRS+Comm_HALO_PACKING_FUSED
#   [r02_e]  tests time:halo time:halo1024 ncu_halo sanitize:halo
#   [r02_f/g] tests:tests/test_algorithm_gpu.py time:sort+halo cub time:halo1024
#   [r02_mgpu, --gpus 2]  smi mgpu_check:2 mgpu_ref:2:16777216:20 mgpu_ref:2:134217728:100:Base_CUDA+Base_B200 \
#                         mgpu_halo:2:512 mgpu_halo:2:1024 mgpu_halo:1:512 mgpu_halo:1:1024
#   [r02_mgpu8, --gpus 8] smi mgpu_check:8 mgpu_ref:4:16777216:20 mgpu_ref:8:16777216:20 \
#                         mgpu_ref:8:134217728:100:Base_CUDA+Base_B200 mgpu_halo:8:512:two mgpu_halo:4:512:two mgpu_halo:8:1024:two mgpu_halo:1:512:two
#   not yet run (the round-2 budget ended): sanitize:reduce+scan+indexlist+pa+sort ; ncu (every default instantiation at HEAD) ;
#                         bench_n:2 bench_n:8 ; mgpu_ref:8:* after the MPI stand-in's ring fix (commit 5b1bbca)
