#!/bin/bash
TAG=${1:-sweep8}
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $(basename $1) $2 n=$3 tally=$4"; MYTRIM_B200_LIB=$PWD/$1 timeout 300 python tools/profile_run.py --workload $2 --primaries $3 --launches 4 --tally $4 2>&1 | tail -2; }
{
for lib in build/variants/*.so; do
  run $lib cu_on_cu_10keV 2097152 1
  run $lib cu_on_cu_10keV 8388608 1
  run $lib c_on_w_1MeV 262144 2
  run $lib xe_on_zro2_500keV 131072 8
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
