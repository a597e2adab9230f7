# Round-2 visit i: ion-log-only clusters variant on the tests/uo2 workload; launches alternating over two engines of one GPU
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/r02i2_overlap.log
: > $LOG
for n in 16384 65536; do
  echo "== clusters-log uo2_fission n=$n" >> $LOG
  timeout 300 python tools/profile_run.py --workload uo2_fission --primaries $n --tally 64 --launches 2 2>&1 | tail -1 >> $LOG
done
timeout 300 python tools/overlap_run.py --workload uo2_fission --primaries 65536 --launches 6 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload uo2_fission --primaries 32768 --launches 8 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload uo2_fission --primaries 32768 --launches 9 --engines 3 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload xe_on_uo2_10MeV --primaries 8192 --tally 1 --launches 8 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload c_on_w_1MeV --primaries 262144 --tally 1 --launches 8 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload xe_on_zro2_500keV --primaries 65536 --tally 1 --launches 8 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload cu_on_cu_10keV --primaries 4194304 --tally 1 --launches 6 >> $LOG 2>&1
cat $LOG
timeout 600 python -m pytest tests -m gpu -x -q -k "uo2 or cluster or per_ion or variant" > $OUT/r02i2_pytest.log 2>&1; tail -5 $OUT/r02i2_pytest.log
