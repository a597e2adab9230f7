#!/bin/bash
# Is the share kernel's deficit a steady-state cost or a tail cost?  Usage: bash tools/sweep4.sh <tag>
TAG=${1:-sweep4}
OUT=gpurun_out
mkdir -p $OUT
run() { # share_below workload primaries [tally]
  echo "== share_below=$1 $2 n=$3 tally=${4:-default}"
  MYTRIM_B200_SHARE_BELOW=$1 timeout 300 python tools/profile_run.py --workload $2 --primaries $3 --launches 3 ${4:+--tally $4} 2>&1 | tail -1
}
{
for n in 1048576 2097152 8388608; do
  run 0 cu_on_cu_10keV $n
  run 1000000 cu_on_cu_10keV $n
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
