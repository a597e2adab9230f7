OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_facade.py -m gpu -q -k "facade_batch or chunks" > $OUT/r02c_pytest.log 2>&1; echo "exit $?" >> $OUT/r02c_pytest.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "coexist" >> $OUT/r02c_pytest.log 2>&1; echo "exit $?" >> $OUT/r02c_pytest.log
{ for v in d4 d6 d8 d6e3; do python tools/compare_libs.py build/variants/base.so build/variants/$v.so; done; python tools/compare_libs.py build/variants/base.so build/variants/d6.so h_on_fe_100keV 20000; } > $OUT/r02c_compare.log 2>&1
bash tools/sweep_variants.sh r02c base d4 d6 d8 d6e3 d8e4 d6m6 base > /dev/null 2>&1
{ echo "== base, 6 CTAs/SM"; MYTRIM_B200_BLOCKS_PER_SM=6 MYTRIM_B200_LIB=$PWD/build/variants/base.so python tools/profile_run.py --primaries 4194304 --launches 4 | tail -2; echo "== base, 5 CTAs/SM"; MYTRIM_B200_BLOCKS_PER_SM=5 MYTRIM_B200_LIB=$PWD/build/variants/base.so python tools/profile_run.py --primaries 4194304 --launches 4 | tail -2; } >> $OUT/r02c_sweep.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 -o $OUT/r02c_prof_uo2 -f python tools/profile_run.py --workload uo2_fission --tally 64 --primaries 4096 --escale 0.05 --launches 2 > $OUT/r02c_prof_uo2.log 2>&1
tail -5 $OUT/r02c_pytest.log; cat $OUT/r02c_compare.log; cat $OUT/r02c_sweep.log; tail -3 $OUT/r02c_prof_uo2.log
