# Round-2 visit h: L1 capacity for the lane stacks (fewer depth bins mirrored in shared memory, explicit carveout);
# launch-size scaling of the tests/uo2 workload.
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/r02h2_l1.log
: > $LOG
run() { # label, env..., -- workload n tally launches
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "== $label $*" >> $LOG
  env "${envs[@]}" timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches $4 2>&1 | tail -2 >> $LOG
}
for hist in default 1024 512 256 0; do
  for carve in default 50 40 100; do
    e=()
    [ $hist != default ] && e+=("MYTRIM_B200_SMEM_HIST=$hist")
    [ $carve != default ] && e+=("MYTRIM_B200_CARVEOUT=$carve")
    run "hist=$hist carve=$carve" "${e[@]}" -- cu_on_cu_10keV 4194304 1 3
  done
done
run "hist=512 h_on_fe" MYTRIM_B200_SMEM_HIST=512 -- h_on_fe_100keV 8388608 1 3
run "hist=default h_on_fe" -- h_on_fe_100keV 8388608 1 3
run "hist=512 xe_zro2" MYTRIM_B200_SMEM_HIST=512 -- xe_on_zro2_500keV 65536 1 3
run "hist=default xe_zro2" -- xe_on_zro2_500keV 65536 1 3
for carve in default 25 50 100; do
  e=(); [ $carve != default ] && e+=("MYTRIM_B200_CARVEOUT=$carve")
  run "uo2 carve=$carve" "${e[@]}" -- uo2_fission 65536 64 2
done
run "uo2 n=262144" -- uo2_fission 262144 64 1
cat $LOG
