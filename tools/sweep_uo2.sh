#!/bin/bash
# A/B timing of library variants on the tests/uo2 workload only.  Usage (under gpurun): bash tools/sweep_uo2.sh <tag> name ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
{ for v in "$@"; do for n in 16384 65536; do echo "== $v uo2_fission n=$n"; MYTRIM_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/profile_run.py --workload uo2_fission --primaries $n --tally 64 --launches 2 2>&1 | tail -1; done; done; } >> $OUT/${TAG}_uo2.log 2>&1
cat $OUT/${TAG}_uo2.log
