# Round-2 visit j: GPU suite, per-configuration rates and bench with the new variants
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/r02j2_pytest.log 2>&1; tail -12 $OUT/r02j2_pytest.log
LOG=$OUT/r02j2_configs.log; : > $LOG
for w in "c_on_w_1MeV 262144 2" "c_on_w_1MeV 262144 1" "xe_on_zro2_500keV 65536 1" "uo2_fission 65536 64"; do
  set -- $w
  echo "== $1 n=$2 tally=$3" >> $LOG
  timeout 300 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches 3 2>&1 | tail -1 >> $LOG
done
cat $LOG
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/r02j2_bench.json 2> $OUT/r02j2_bench.err; tail -3 $OUT/r02j2_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j2_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d["configs"].items():
    print(k, "%.3e %.3e" % (v["cascades_per_s"], v["collision_steps_per_s"]), v.get("two_engines_per_gpu", {}).get("collision_steps_per_s"))
PY
cd /tmp && MYTRIM_TIMING=1 MYTRIM_SEED=39172 timeout 300 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 appout 10 0.1 131072 2>&1 | tail -4
MYTRIM_ENGINES_PER_GPU=1 MYTRIM_TIMING=1 MYTRIM_SEED=39172 timeout 300 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 appout1 10 0.1 131072 2>&1 | tail -4
cmp appout.Erec appout1.Erec && cmp appout.dist appout1.dist && echo "files identical for 1 and 2 engines per GPU"
