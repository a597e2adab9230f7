#!/bin/bash
TAG=${1:-sweep7}
OUT=gpurun_out
mkdir -p $OUT
run() { # lib workload primaries tally
  echo "== $(basename $1) $2 n=$3 tally=$4"
  MYTRIM_B200_LIB=$PWD/$1 timeout 300 python tools/profile_run.py --workload $2 --primaries $3 --launches 3 --tally $4 2>&1 | tail -1
}
{
for lib in build/variants/*.so; do
  run $lib cu_on_cu_10keV 2097152 4
  run $lib cu_on_cu_10keV 2097152 8
  run $lib c_on_w_1MeV 65536 8
  run $lib xe_on_zro2_500keV 32768 8
  run $lib uo2_fission 8192 0
  run $lib uo2_fission 65536 0
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
