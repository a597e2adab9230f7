OUT=$GRAFT_REPO_ROOT/gpurun_out; LOG=$OUT/r02k3.log; : > $LOG
cd /tmp
for e in 1 2 3; do
  echo "== mytrim_uo2 196608 events (six chunks), engines per GPU $e" >> $LOG
  ( time MYTRIM_ENGINES_PER_GPU=$e MYTRIM_TIMING=1 MYTRIM_SEED=39172 timeout 400 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 app$e 10 0.1 196608 ) >> $LOG 2>&1
  echo "rc=$?" >> $LOG
done
cmp app1.Erec app3.Erec && cmp app1.dist app2.dist && echo "files identical for 1, 2 and 3 engines per GPU" >> $LOG
cat $LOG
