#!/bin/bash
# A/B timing of library variants on the workloads that run through the work-sharing kernels (fewer primaries than 4 per
# lane): the tests/uo2 fission fragments, Xe->ZrO2 500 keV, C->W 1 MeV.  Usage (under gpurun): bash tools/sweep_share_workloads.sh <tag> name ...
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
{
for v in "$@"; do
  for w in "uo2_fission 16384 64" "uo2_fission 65536 64" "xe_on_zro2_500keV 65536 1" "c_on_w_1MeV 262144 2"; do
    set -- $w
    echo "== $v $1 n=$2 tally=$3 ${MYTRIM_B200_SHARE_MIN_E:+minE=$MYTRIM_B200_SHARE_MIN_E}"
    MYTRIM_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches 2 2>&1 | tail -1
  done
done
} >> $OUT/${TAG}_share.log 2>&1
cat $OUT/${TAG}_share.log
