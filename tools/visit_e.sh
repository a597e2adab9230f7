OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/r02e_pytest.log 2>&1; echo "exit $?" >> $OUT/r02e_pytest.log
bash tools/sweep_variants.sh r02e base2 cur nr9 base2 > /dev/null 2>&1
MYTRIM_B200_NO_NOREC=1 python tools/compare_libs.py build/variants/base2.so build/variants/cur.so > $OUT/r02e_compare.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/r02e_bench.json 2> $OUT/r02e_bench.err; echo "bench exit $?" >> $OUT/r02e_bench.err
timeout 500 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section LaunchStats --section Occupancy --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 -o $OUT/r02e_prof_uo2 -f python tools/profile_run.py --workload uo2_fission --tally 64 --ionlog-z 54 --primaries 8192 --launches 2 > $OUT/r02e_prof_uo2.log 2>&1
tail -25 $OUT/r02e_pytest.log; grep "^==\|launch 3" $OUT/r02e_sweep.log; cat $OUT/r02e_compare.log; cat $OUT/r02e_bench.json | cut -c1-600; tail -3 $OUT/r02e_prof_uo2.log
