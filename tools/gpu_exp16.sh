#!/bin/bash
H=rajaperf_b200/suite/raja-perf-b200.exe
for args in "--checkrun 20 --graph" "--checkrun 100 --graph" "--checkrun 20" "--checkrun 100"; do
  $H -k REDUCE_SUM DOT --size 134217728 $args --outdir /tmp/rs > /dev/null 2>&1
  echo "== $args"; cut -d, -f1,4,8,9 /tmp/rs/RAJAPerf-bandwidth.csv | tail -2
done
