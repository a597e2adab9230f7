#!/bin/bash
# Builds library variants (launch-bounds experiments) into build/variants for tools/sweep*.sh.
# Usage: bash tools/build_variants.sh name "-DMACRO=.. -DMACRO=.." [name flags ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  nvcc $2 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -I include \
    -o build/variants/$1.so mytrim_b200/csrc/mtb_engine.cu mytrim_b200/csrc/facade.cpp
  echo "built build/variants/$1.so ($2)"
  shift 2
done
