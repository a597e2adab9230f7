#!/bin/bash
# Work-sharing threshold: time launches of several sizes with sharing forced on and off.
# Usage (under gpurun): bash tools/sweep_share.sh <tag>
TAG=${1:-share}
OUT=gpurun_out
mkdir -p $OUT
run() { # share_below workload primaries
  echo "== share_below=$1 $2 n=$3"
  MYTRIM_B200_SHARE_BELOW=$1 timeout 300 python tools/profile_run.py --workload $2 --primaries $3 --launches 3 2>&1 | tail -1
}
{
for n in 65536 131072 262144 524288 1048576 2097152; do
  run 0 cu_on_cu_10keV $n
  run 1000000 cu_on_cu_10keV $n
done
for n in 4096 16384 65536 262144; do
  run 0 c_on_w_1MeV $n
  run 1000000 c_on_w_1MeV $n
done
for n in 2048 8192 32768 131072; do
  run 0 xe_on_zro2_500keV $n
  run 1000000 xe_on_zro2_500keV $n
done
for n in 65536 262144 1048576 4194304; do
  run 0 h_on_fe_100keV $n
  run 1000000 h_on_fe_100keV $n
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
