OUT=$GRAFT_REPO_ROOT/gpurun_out; LOG=$OUT/r02m2_phases.log; : > $LOG
cd /tmp
for e in 1 2; do
echo "== 262144 events, engines per GPU $e" >> $LOG
( time MYTRIM_ENGINES_PER_GPU=$e MYTRIM_TIMING=1 MYTRIM_SEED=39172 timeout 400 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 p$e 10 0.1 262144 ) 2>&1 | grep "workload\|ERROR\|engine\|real\|user" >> $LOG
done
cat $LOG
