# Round-2 visit l: device-wide work-sharing ring — result neutrality, then timings on/off and donation thresholds
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/r02l_gpool.log; : > $LOG
timeout 120 python tools/profile_run.py --workload xe_on_uo2_10MeV --primaries 512 --tally 1 --launches 2 >> $LOG 2>&1 || { echo "SMALL RUN FAILED rc=$?" >> $LOG; cat $LOG; exit 1; }
timeout 300 python -m pytest tests -m gpu -x -q -k "sharing or share or lean_variants or uo2" >> $LOG 2>&1
run() { # label env... -- workload n tally launches
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "== $label $*" >> $LOG
  env "${envs[@]}" timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches $4 2>&1 | tail -1 >> $LOG
}
for w in "xe_on_uo2_10MeV 8192 1 3" "uo2_fission 65536 64 2" "uo2_fission 16384 64 2" "c_on_w_1MeV 262144 2 3" "xe_on_zro2_500keV 65536 1 3" "cu_on_cu_150keV 131072 1 3"; do
  run "off" MYTRIM_B200_NO_GLOBAL_SHARE=1 -- $w
  run "on(2000eV)" -- $w
  run "on(500eV)" MYTRIM_B200_GSHARE_MIN_E=500 -- $w
  run "on(10keV)" MYTRIM_B200_GSHARE_MIN_E=10000 -- $w
done
timeout 300 python tools/overlap_run.py --workload uo2_fission --primaries 65536 --launches 4 >> $LOG 2>&1
timeout 300 python tools/overlap_run.py --workload xe_on_uo2_10MeV --primaries 8192 --tally 1 --launches 8 >> $LOG 2>&1
cat $LOG
