#!/bin/bash
# One short GPU-box visit (tight per-step limits): smoke, GPU parity tests, bench (both arms), ncu full capture + launch list.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag] [skip-list: smoke,pytest,bench,ref,ncu,launches,configs]
TAG=${1:-r02}
SKIP=",${2:-},"
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
want() { [[ "$SKIP" != *",$1,"* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
if want smoke; then
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_smoke.log
fi
if want pytest; then
timeout 600 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_pytest_gpu.log
fi
if want bench; then
timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_bench.err
fi
if want ref; then
timeout 180 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "ref exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_bench_ref.err
fi
if want ncu; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o $OUT/${TAG}_prof -f python tools/profile_run.py --primaries 2097152 --launches 2 > $OUT/${TAG}_prof.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_prof.log
fi
if want launches; then
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
echo "ncu launches exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_launches.log
fi
if want configs; then
bash tools/bench_configs.sh ${TAG}_configs > /dev/null 2>&1
echo "configs exit $? at $(( $(date +%s) - T0 )) s" >> $OUT/${TAG}_configs.log
fi
tail -3 $OUT/${TAG}_smoke.log; tail -14 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_ref.json; tail -2 $OUT/${TAG}_prof.log
