#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench (both arms), ncu launch list + full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_run.py --primaries 2097152 --launches 3 > $OUT/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_prof -f python tools/profile_run.py --primaries 2097152 --launches 2 > $OUT/${TAG}_prof.log 2>&1
tail -3 $OUT/${TAG}_smoke.log; tail -5 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_ref.json
