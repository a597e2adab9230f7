#!/bin/bash
TAG=${1:-sweep5}
OUT=gpurun_out
mkdir -p $OUT
{
for lib in build/variants/*.so; do
  for sb in 0 1000000; do
    echo "== $(basename $lib) share_below=$sb"
    MYTRIM_B200_LIB=$PWD/$lib MYTRIM_B200_SHARE_BELOW=$sb timeout 300 python tools/profile_run.py --workload cu_on_cu_10keV --primaries 2097152 --launches 3 2>&1 | tail -1
  done
done
} > $OUT/${TAG}.log 2>&1
cat $OUT/${TAG}.log
