# Round-2 visit k: FAST-PHONON variant, engines per GPU in apps/mytrim_uo2, variant tests
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/r02k2.log; : > $LOG
timeout 300 python -m pytest tests -m gpu -x -q -k "lean_variants or follow_policies or phonon" >> $LOG 2>&1
for w in "xe_on_zro2_500keV 131072 8" "xe_on_zro2_500keV 65536 8" "cu_on_cu_10keV 4194304 8"; do
  set -- $w
  echo "== $1 n=$2 tally=$3" >> $LOG
  timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches 3 2>&1 | tail -1 >> $LOG
done
cd /tmp
for e in 1 2 3; do
  echo "== mytrim_uo2 196608 events (six chunks), engines per GPU $e" >> $GRAFT_REPO_ROOT/$LOG
  MYTRIM_ENGINES_PER_GPU=$e MYTRIM_TIMING=1 MYTRIM_SEED=39172 timeout 400 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 app$e 10 0.1 196608 2>&1 | grep workload >> $GRAFT_REPO_ROOT/$LOG
done
cmp app1.Erec app3.Erec && cmp app1.dist app2.dist && echo "files identical for 1, 2 and 3 engines per GPU" >> $GRAFT_REPO_ROOT/$LOG
cat $GRAFT_REPO_ROOT/$LOG
