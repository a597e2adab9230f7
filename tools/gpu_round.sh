#!/bin/bash
# One short GPU-box visit (tight per-step limits): smoke, GPU parity tests, bench (both arms), ncu full capture + launch list.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_smoke.log
timeout 480 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_pytest_gpu.log
timeout 240 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_bench.err
timeout 180 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "ref exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_prof -f python tools/profile_run.py --primaries 2097152 --launches 2 > $OUT/${TAG}_prof.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_prof.log
timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_run.py --primaries 2097152 --launches 3 > $OUT/${TAG}_launches.log 2>&1
echo "ncu launches exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_launches.log
tail -3 $OUT/${TAG}_smoke.log; tail -14 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_ref.json; tail -2 $OUT/${TAG}_prof.log
