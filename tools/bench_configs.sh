#!/bin/bash
# Throughput of every BASELINE.json configuration (resident launches, default engine settings) plus the
# batched mytrim_uo2 driver.  Usage (under gpurun, from the repo root): bash tools/bench_configs.sh <tag>
TAG=${1:-configs}
OUT=$PWD/gpurun_out
mkdir -p $OUT
run() { # workload primaries [tally]
  echo "== $1 n=$2 tally=${3:-1}"
  timeout 600 python tools/profile_run.py --workload $1 --primaries $2 --launches 3 ${3:+--tally $3} 2>&1 | tail -1
}
{
run cu_on_cu_10keV 8388608
run cu_on_cu_1keV 16777216
run h_on_fe_100keV 16777216
run he_on_fe_100keV 2097152
run c_on_w_1MeV 262144
run c_on_w_1MeV 262144 2     # vacenergycount, as validation/c_on_w/input.json
run xe_on_zro2_500keV 131072
run xe_on_zro2_500keV 131072 8
echo "== mytrim_uo2 (apps/mytrim_uo2.cpp, 32768 fission events = 65536 fragments per launch)"
(cd /tmp && MYTRIM_SEED=39172 MYTRIM_TIMING=1 timeout 900 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 uo2out 10 0.1 32768 2>&1 | grep workload)
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
