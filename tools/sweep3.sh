#!/bin/bash
# Share-kernel register budgets: variants x (share on/off) x workloads.  Usage: bash tools/sweep3.sh <tag>
TAG=${1:-sweep3}
OUT=gpurun_out
mkdir -p $OUT
run() { # lib share_below workload primaries [tally]
  echo "== $(basename $1) share_below=$2 $3 n=$4 tally=${5:-default}"
  MYTRIM_B200_LIB=$PWD/$1 MYTRIM_B200_SHARE_BELOW=$2 timeout 300 python tools/profile_run.py --workload $3 --primaries $4 --launches 3 ${5:+--tally $5} 2>&1 | tail -1
}
{
for lib in build/variants/*.so; do
  run $lib 0 cu_on_cu_10keV 2097152
  run $lib 1000000 cu_on_cu_10keV 2097152
  run $lib 1000000 cu_on_cu_10keV 262144
  run $lib 0 cu_on_cu_10keV 2097152 8
  run $lib 1000000 cu_on_cu_10keV 2097152 8
  run $lib 1000000 c_on_w_1MeV 65536
  run $lib 1000000 c_on_w_1MeV 65536 8
  run $lib 1000000 xe_on_zro2_500keV 32768
  run $lib 0 h_on_fe_100keV 4194304
  run $lib 1000000 h_on_fe_100keV 4194304
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
