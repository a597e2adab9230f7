#!/bin/bash
# A/B timing of library variants on one launch size of every BASELINE configuration.  Usage: bash tools/sweep_all_configs.sh <tag> name ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
{ for v in "$@"; do
  for w in "cu_on_cu_10keV 4194304 1" "h_on_fe_100keV 8388608 1" "he_on_fe_100keV 2097152 1" "c_on_w_1MeV 262144 2" "xe_on_zro2_500keV 65536 1" "uo2_fission 65536 64"; do
    set -- $w
    echo "== $v $1 n=$2 tally=$3"
    MYTRIM_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches 3 2>&1 | tail -1
  done
done; } >> $OUT/${TAG}_all.log 2>&1
cat $OUT/${TAG}_all.log
