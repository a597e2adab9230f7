#!/bin/bash
# Runs tools/profile_run.py against every library variant under build/variants (launch bounds sweep).
OUT=gpurun_out
mkdir -p $OUT
for lib in build/variants/*.so; do
  echo "== $lib"
  MYTRIM_B200_LIB=$PWD/$lib timeout 120 python tools/profile_run.py --primaries ${1:-1048576} --launches 4 2>&1 | tail -2
done > $OUT/${2:-sweep}.log 2>&1
cat $OUT/${2:-sweep}.log
