OUT=gpurun_out; mkdir -p $OUT
# unmodified reference apps/mytrim_uo2.C: through the façade on the GPU (TrimBase::trim + host hook replay) and the reference binary on one core
( cd /tmp; export MYTRIM_SEED=39172 MYTRIM_DATADIR=$GRAFT_REPO_ROOT/oracle/_ref/data
  for n in 1 4; do
    /usr/bin/time -f "facade_app events=$n wall=%e s" $GRAFT_REPO_ROOT/oracle/_ref/facade_apps/mytrim_uo2 fa$n 10 0.1 $n > fa$n.out 2> fa$n.err; tail -1 fa$n.err
    /usr/bin/time -f "reference_binary events=$n wall=%e s" $GRAFT_REPO_ROOT/oracle/_ref/mytrim_uo2 rf$n 10 0.1 $n > rf$n.out 2> rf$n.err; tail -1 rf$n.err
    wc -l fa$n.Erec rf$n.Erec | head -2
  done
  /usr/bin/time -f "batched_driver events=4 wall=%e s" $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 ba4 10 0.1 4 > ba4.out 2> ba4.err; tail -1 ba4.err
) > $OUT/r02f_facade_timing.log 2>&1
cat $OUT/r02f_facade_timing.log
bash tools/sanitize.sh r02 > $OUT/r02f_sanitize_summary.log 2>&1
cat $OUT/r02f_sanitize_summary.log
