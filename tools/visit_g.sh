OUT=gpurun_out; mkdir -p $OUT
python - > $OUT/r02g_facade_timing.log 2>&1 <<'PY'
import os, subprocess, time
root = os.environ.get("GRAFT_REPO_ROOT", os.getcwd())
env = dict(os.environ, MYTRIM_SEED="39172", MYTRIM_DATADIR=os.path.join(root, "oracle/_ref/data"))
os.chdir("/tmp")
def run(tag, exe, n):
    t = time.perf_counter()
    p = subprocess.run([exe, tag, "10", "0.1", str(n)], env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t
    lines = sum(1 for _ in open(tag + ".Erec")) if os.path.exists(tag + ".Erec") else -1
    print("%-34s events=%d wall=%.2f s rc=%d Erec lines=%d" % (tag, n, dt, p.returncode, lines), flush=True)
for n in (1, 4):
    run("facade_app_unmodified_mytrim_uo2_C", os.path.join(root, "oracle/_ref/facade_apps/mytrim_uo2"), n)
    run("reference_binary_one_core", os.path.join(root, "oracle/_ref/mytrim_uo2"), n)
run("batched_driver_apps_mytrim_uo2_cpp", os.path.join(root, "build/apps/mytrim_uo2"), 4)
run("batched_driver_apps_mytrim_uo2_cpp", os.path.join(root, "build/apps/mytrim_uo2"), 1024)
PY
cat $OUT/r02g_facade_timing.log
sed -i 's/for tool in memcheck racecheck initcheck synccheck; do/for tool in ${SANITIZE_TOOLS:-memcheck racecheck initcheck synccheck}; do/' tools/sanitize.sh
SANITIZE_TOOLS=initcheck bash tools/sanitize.sh r02 | tail -3
