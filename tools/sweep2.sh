#!/bin/bash
# A/B timing of library variants under build/variants on the BASELINE workloads (resident launches).
# Usage (under gpurun, from the repo root): bash tools/sweep2.sh <tag>
TAG=${1:-sweep2}
OUT=gpurun_out
mkdir -p $OUT
run() { # lib bps workload primaries tally
  echo "== $(basename $1) bps=$2 $3 n=$4"
  MYTRIM_B200_LIB=$PWD/$1 MYTRIM_B200_BLOCKS_PER_SM=$2 timeout 300 python tools/profile_run.py --workload $3 --primaries $4 --launches 3 ${5:+--tally $5} 2>&1 | tail -2
}
{
for lib in build/variants/*.so; do
  for bps in 6 7 8; do
    run $lib $bps cu_on_cu_10keV 2097152
  done
done
for lib in build/variants/*.so; do
  run $lib 8 cu_on_cu_1keV 4194304
  run $lib 8 h_on_fe_100keV 1048576
  run $lib 8 he_on_fe_100keV 262144
  run $lib 8 c_on_w_1MeV 65536
  run $lib 8 xe_on_zro2_500keV 32768
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
