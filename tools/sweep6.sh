#!/bin/bash
# Donation threshold of the work-sharing pool.  Usage: bash tools/sweep6.sh <tag>
TAG=${1:-sweep6}
OUT=gpurun_out
mkdir -p $OUT
run() { # min_E workload primaries
  echo "== min_E=$1 $2 n=$3"
  MYTRIM_B200_SHARE_BELOW=1000000 MYTRIM_B200_SHARE_MIN_E=$1 timeout 300 python tools/profile_run.py --workload $2 --primaries $3 --launches 3 2>&1 | tail -1
}
{
run 100 cu_on_cu_10keV 2097152
for e in 0 100 1000 10000; do
  run $e cu_on_cu_10keV 65536
  run $e c_on_w_1MeV 4096
  run $e c_on_w_1MeV 65536
  run $e xe_on_zro2_500keV 2048
  run $e xe_on_zro2_500keV 32768
  run $e h_on_fe_100keV 65536
  run $e h_on_fe_100keV 262144
done
cd /tmp
for e in 0 100 1000 10000; do
  echo "== uo2 chunk 2048 min_E=$e"
  MYTRIM_B200_SHARE_MIN_E=$e MYTRIM_SEED=39172 MYTRIM_TIMING=1 timeout 600 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 uo2out 10 0.1 8192 2>&1 | grep workload
done
echo "== uo2 chunk 32768 min_E=100"
MYTRIM_UO2_CHUNK=32768 MYTRIM_B200_SHARE_MIN_E=100 MYTRIM_SEED=39172 MYTRIM_TIMING=1 timeout 600 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 uo2out 10 0.1 32768 2>&1 | grep workload
} > $GRAFT_REPO_ROOT/$OUT/${TAG}.log 2>&1
cat $GRAFT_REPO_ROOT/$OUT/${TAG}.log
