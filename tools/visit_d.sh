OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/r02d_pytest.log 2>&1; echo "exit $?" >> $OUT/r02d_pytest.log
bash tools/sweep_variants.sh r02d base2 norec norec8 base2 > /dev/null 2>&1
python tools/compare_libs.py build/variants/base2.so build/variants/norec.so > $OUT/r02d_compare.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 -o $OUT/r02d_prof_uo2 -f python tools/profile_run.py --workload uo2_fission --tally 64 --ionlog-z 54 --primaries 16384 --launches 2 > $OUT/r02d_prof_uo2.log 2>&1
tail -25 $OUT/r02d_pytest.log; grep "^==\|launch 3" $OUT/r02d_sweep.log; cat $OUT/r02d_compare.log; tail -3 $OUT/r02d_prof_uo2.log
