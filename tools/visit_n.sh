# Round-2 visit n: ring size, sharing threshold and work sharing on/off on the few-primaries workloads (final library)
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/r02n3_share.log; : > $LOG
run() { # label env... -- workload n tally launches
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "== $label $*" >> $LOG
  env "${envs[@]}" timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches $4 2>&1 | tail -1 >> $LOG
}
for v in cur12 p16 p64; do
  for w in "uo2_fission 65536 64 2" "xe_on_uo2_10MeV 8192 1 3" "c_on_w_1MeV 262144 2 3"; do
    run "ring $v" MYTRIM_B200_LIB=$PWD/build/variants/$v.so -- $w
  done
done
for e in 150 600 1200; do
  run "min_E $e" MYTRIM_B200_SHARE_MIN_E=$e -- uo2_fission 65536 64 2
  run "min_E $e" MYTRIM_B200_SHARE_MIN_E=$e -- xe_on_uo2_10MeV 8192 1 3
done
for w in "c_on_w_1MeV 262144 2 3" "cu_on_cu_150keV 131072 1 3" "xe_on_zro2_500keV 131072 1 3" "xe_on_zro2_500keV_x 0 0 0"; do
  set -- $w; [ "$2" = 0 ] && continue
  run "no share" MYTRIM_B200_NO_SHARE=1 -- $w
  run "share" -- $w
done
cat $LOG
