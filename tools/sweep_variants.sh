#!/bin/bash
# A/B timing of library variants built by tools/build_variants.sh (or by hand into build/variants/):
# resident launches of the north-star workload through each variant, first and last the baseline.
# Usage (under gpurun, from the repo root): bash tools/sweep_variants.sh <tag> name [name ...]
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
{
for v in "$@"; do
  echo "== $v"
  MYTRIM_B200_LIB=$PWD/build/variants/$v.so timeout 120 python tools/profile_run.py --primaries 4194304 --launches 4 2>&1 | tail -3
done
} > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
