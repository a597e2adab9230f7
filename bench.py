#!/usr/bin/env python3
"""bench.py — full ion cascades per second.

Headline workload (BASELINE.json / SURVEY.md §8d config 1): Cu (Z=29, m=63.546) ions at 10 keV into a
1000 A Cu layer (rho 8.92), full recoil cascades (follow ALL), TrimVacCount tallies.  A "step" is one batch
of `--primaries` cascades per GPU through the transport kernel.  `--workload` times one of the other
BASELINE.json configurations instead (h_on_fe_100keV, he_on_fe_100keV, c_on_w_1MeV, xe_on_zro2_500keV,
uo2_fission); the default run also measures every one of them once per GPU and reports them under
"configs" (the headline metric and config do not change).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

value   = cascades/s with the primaries already resident in HBM (device time of the K launches, CUDA events on
          the launching stream, max over ranks, plus the ONE tally join of the job when N > 1)
e2e     = the same metric through mtb_run() with HOST buffers: every step copies its primaries host->device
          (pinned memory) and reads the tallies back; the tally join over the ranks is inside the clock.
The reference arm (--impl reference) times the UNMODIFIED reference library (oracle/_ref) on all host cores on a
bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_STEP = 930.0  # FP32-equivalent flops per collision step, SURVEY.md §8d
HEADLINE = "cu_on_cu_10keV"
MASTER_SEED = 2344
FP32_LANES = 148 * 128  # FP32 lanes of a B200: 148 SMs x 4 sub-partitions x 32
# configurations whose launches have few primaries per lane (work-sharing kernels): also timed with two engines per GPU
PIPELINED_CONFIGS = ("c_on_w_1MeV", "xe_on_zro2_500keV", "xe_on_uo2_10MeV", "uo2_fission")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference library on the host cores
# ---------------------------------------------------------------------------------------------------------
REF_TALLY = {1: "vaccount", 2: "vacenergycount"}


def reference_cascades_per_s(workload, n, threads, timeout=900):
    """Times oracle/_ref/ref_driver (the unmodified reference library, runmytrim-equivalent set-up)."""
    from mytrim_b200 import workloads
    from tests import util
    wl = workloads.BENCH_WORKLOADS[workload]
    c = workloads.CONFIGS[wl.get("config", workload)]
    tally = wl.get("ref_tally") or REF_TALLY.get(wl["tally"], "vaccount")
    lines = util.reference_script(c["ion"], c["materials"], c["thicknesses"], n=n, tally=tally,
                                  threads=threads, master=MASTER_SEED, box=c.get("box"))
    lines.append("run")
    out = util.run_reference("\n".join(lines) + "\n", timeout=timeout)
    return json.loads(out[-1])


def reference_uo2_primaries_per_s(events_per_proc, procs, timeout=900):
    """The reference's own apps/mytrim_uo2.C (oracle/_ref/mytrim_uo2, single-threaded by design): one process per
    core, each with its own seed and `events_per_proc` fission events of the gold geometry (r = 10, Cbf = 0.1)."""
    import tempfile
    from tests import util
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        ps = []
        for k in range(procs):
            env = util.ref_env()
            env["MYTRIM_SEED"] = str(39172 + k)
            ps.append(subprocess.Popen([util.REF_UO2, os.path.join(tmp, "o%d" % k), "10", "0.1", str(events_per_proc)],
                                       env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        for p in ps:
            p.wait(timeout=timeout)
    dt = time.perf_counter() - t0
    n = 2 * events_per_proc * procs
    return {"n": n, "seconds": dt, "cascades_per_s": n / dt, "steps": 0}


# reference cascades per core and step (a step of the reference arm is ~2-20 s of wall time on all cores)
REF_PER_CORE = {"cu_on_cu_10keV": 400, "h_on_fe_100keV": 2000, "he_on_fe_100keV": 100, "c_on_w_1MeV": 8,
                "xe_on_zro2_500keV": 2, "xe_on_zro2_500keV_trimrecoils": 40, "uo2_fission": 2, "cu_on_cu_150keV": 20, "h_on_fe_1MeV": 1000, "xe_on_uo2_10MeV": 1}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build_test_infrastructure()
    from mytrim_b200 import workloads
    from tests import util
    cores = os.cpu_count() or 1
    name = args.workload
    if not util.have_reference() or (name == "uo2_fission" and not os.path.exists(util.REF_UO2)):
        emit({"impl": "reference", "unavailable": "oracle/_ref was not built (reference tree absent)"})
        return 0
    per_step = args.ref_cascades if args.ref_cascades else REF_PER_CORE[name] * cores

    def once(n):
        if name == "uo2_fission":
            return reference_uo2_primaries_per_s(max(1, n // (2 * cores)), cores)
        return reference_cascades_per_s(name, n, cores)

    for _ in range(args.warmup):
        once(max(per_step // 8, cores))
    t_total, n_total, steps_total = 0.0, 0, 0
    for _ in range(args.steps):
        r = once(per_step)
        t_total += r["seconds"]
        n_total += r["n"]
        steps_total += r["steps"]
    value = n_total / t_total
    sample = "%d steps x %d cascades of %s on %d threads (unmodified reference)" % (args.steps, n_total // args.steps,
                                                                                 name, cores)
    line = {
        "impl": "reference", "metric": "cascades_per_s", "value": value, "unit": "cascades/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "description": workloads.BENCH_WORKLOADS[name]["desc"],
                   "primaries_per_step": n_total // args.steps},
        "collision_steps_per_s": steps_total / t_total if steps_total else None,
        "cpu_baseline": {"value": value, "unit": "cascades/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "cascades/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version on
    communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the result
    line goes to a private duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(line) + "\n").encode())


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def c_abi_allreduce_check(timeout=120):
    """mtb_allreduce (the C-ABI tally join of a single-process multi-GPU job, dlopen'd NCCL) on GPUs 0 and 1 in a
    child process with a hard time limit: two handles, primaries sharded by global index, reduced tallies equal
    one GPU running everything."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "allreduce_check.py")], capture_output=True,
                           text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return "timeout"
    last = p.stdout.strip().split("\n")[-1] if p.stdout.strip() else ""
    if p.returncode == 0 and (last.startswith("ok") or last.startswith("skipped")):
        return last
    return "failed: " + (last or p.stderr.strip()[-200:])


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, help="BASELINE.json configuration to time")
    ap.add_argument("--primaries", type=int, default=0, help="cascades per GPU per step (default: per workload)")
    ap.add_argument("--ref-cascades", type=int, default=0, help="cascades per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the one-launch-per-configuration sweep")
    args = ap.parse_args()

    from mytrim_b200 import workloads
    if args.workload not in workloads.BENCH_WORKLOADS:
        raise SystemExit("unknown workload %s (have: %s)" % (args.workload, ", ".join(workloads.BENCH_WORKLOADS)))
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the transport engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        g.build_engine()
    if world > 1:
        dist.barrier()

    from mytrim_b200 import capi
    from mytrim_b200 import dist as mdist
    lib = capi.load_library()

    name = args.workload
    wl = workloads.BENCH_WORKLOADS[name]
    B = args.primaries if args.primaries else wl["primaries"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def make_engine(wname, n):
        w = workloads.BENCH_WORKLOADS[wname]
        kw = dict(tally_mask=w["tally"], device=local_rank)
        kw.update(w.get("engine", {}))
        if w["tally"] & capi.TALLY_IONLOG:
            kw.update(ionlog_z=w.get("ionlog_z", 0), ionlog_capacity=max(1 << 20, 64 * n))
        e = capi.Engine(**kw)
        # primaries in pinned host memory (the e2e arm copies them every step); every rank its own share of the
        # fission events, the beams are identical primaries
        host = torch.empty(n * capi.ION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        arr = np.frombuffer(host.numpy(), dtype=capi.ION_DTYPE)
        arr[:] = workloads.setup_workload(e, wname, n, first_primary=rank * n)
        return e, host

    eng, pinned = make_engine(name, B)
    variants = {name: eng.kernel_variant()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the tally join: ONE all-gather of the engine's tally blocks + a local reduction (mytrim_b200/dist.py),
    # the analogue of runmytrim's threadJoin
    eng.upload_primaries_ptr(B, pinned.data_ptr())
    eng.synchronize()
    t_u64, t_f64 = mdist.tally_tensors(eng)
    reducer = mdist.TallyReducer(t_u64, t_f64)

    def reduce_tallies(write_back=True):
        if world == 1:
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reducer.reduce(write_back=write_back)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def first_index(step):
        return (step * world + rank) * B

    has_log = bool(wl["tally"] & capi.TALLY_IONLOG)

    # ---------------- resident arm: `value` ----------------
    step_id = 0
    for _ in range(args.warmup):
        eng.launch_resident(MASTER_SEED, first_index(step_id))
        eng.synchronize()
        if has_log:
            lib.mtb_clear_lists(eng._h)
        step_id += 1
    for _ in range(args.warmup):
        reduce_tallies(write_back=False)  # warm-up of the collective on the very buffers the timed join uses
    eng.reset_tallies()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the CUDA-event brackets)
        torch.cuda.synchronize()
        eng.launch_resident(MASTER_SEED, first_index(step_id))
        eng.synchronize()
        dev_ms += eng.last_kernel_ms()
        if has_log:
            lib.mtb_clear_lists(eng._h)
        step_id += 1
    barrier()
    wall_resident = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    counters = eng.counters()  # this rank's tallies over the timed steps
    # All ranks enter the join together: stopping the clock sampler (a subprocess) and reading the counters takes
    # rank-dependent tens of milliseconds of host time, which a rank that arrives early would otherwise sit out inside
    # the collective and report as "reduction time" (round 1: 50 ms at N = 2 and 4 against 0.25 ms at N = 8).
    barrier()
    red_ms = reduce_tallies()  # whole-job tallies (one join per job, as in runmytrim's threadJoin)
    total = eng.counters() if world > 1 else counters
    reduction_check = None
    if world > 1:
        # self-check of the join: the totals every rank now holds equal the sum of the W contributions
        pu, pf = reducer.per_rank()
        ok = bool((pu[:, :mdist.N_ADDITIVE_COUNTERS].sum(dim=0) == t_u64[:mdist.N_ADDITIVE_COUNTERS]).all()) and \
            bool((pu[:, mdist.N_COUNTER_SLOTS:].sum(dim=0) == t_u64[mdist.N_COUNTER_SLOTS:]).all()) and \
            bool(torch.allclose(pf.sum(dim=0), t_f64, rtol=1e-12)) and \
            int(t_u64[4]) == B * world * args.steps
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        reduction_check = "ok" if int(flag[0]) == 1 else "MISMATCH"
    t_rank = torch.tensor([dev_ms + red_ms, dev_ms, float(counters["steps"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t_rank.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_rank.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax, tsum = t_rank, t_rank
    job_ms = float(tmax[0])
    kernel_ms = float(tmax[1])
    coll_steps = float(tsum[2])
    cascades = float(B) * world * args.steps
    value = cascades / (job_ms * 1e-3)

    # ---------------- end-to-end arm: host buffers through mtb_run ----------------
    vac_host = np.zeros(1 << 14, dtype=np.uint64)
    repl_host = np.zeros(1 << 14, dtype=np.uint64)
    evac_host = np.zeros((32, 1 << 14), dtype=np.uint64) if wl["tally"] & capi.TALLY_VAC_ENERGY else None
    log_host = np.zeros(max(1 << 20, 64 * B), dtype=capi.IONLOG_DTYPE) if has_log else None
    nb = ctypes.c_size_t()
    cnt = capi.Counters()
    d2h = [0]

    def e2e_step(step):
        rc = lib.mtb_run(eng._h, B, pinned.data_ptr(), MASTER_SEED, first_index(step), None)
        if rc != 0:
            raise RuntimeError(lib.mtb_last_error().decode())
        lib.mtb_get_counters(eng._h, ctypes.byref(cnt))
        nbytes = ctypes.sizeof(capi.Counters)
        if wl["tally"] & capi.TALLY_VAC_DEPTH:
            lib.mtb_get_vac_depth(eng._h, vac_host.ctypes.data, repl_host.ctypes.data, len(vac_host), ctypes.byref(nb))
            nbytes += vac_host.nbytes + repl_host.nbytes
        if evac_host is not None:
            lib.mtb_get_vac_energy(eng._h, evac_host.ctypes.data, evac_host.shape[0], evac_host.shape[1])
            nbytes += evac_host.nbytes
        if has_log:
            lib.mtb_get_ion_log(eng._h, log_host.ctypes.data, len(log_host), ctypes.byref(nb))
            lib.mtb_clear_lists(eng._h)
            nbytes += nb.value * capi.IONLOG_DTYPE.itemsize
        d2h[0] = nbytes

    e2e_step(step_id)
    step_id += 1
    eng.reset_tallies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step(step_id)
        step_id += 1
    if world > 1:
        reducer.reduce(write_back=True)  # the job's tally join belongs to the end-to-end time
        lib.mtb_get_counters(eng._h, ctypes.byref(cnt))
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = cascades / float(t_e[0])
    e2e_primaries_total = int(cnt.primaries)
    eng.close()

    # ---------------- every BASELINE.json configuration once per GPU ----------------
    configs = {}
    if not args.no_configs:
        for cname, cw in workloads.BENCH_WORKLOADS.items():
            n = cw["primaries"]
            if cname == name:
                ms, st = kernel_ms / args.steps, counters["steps"] / args.steps
                n = B
            else:
                e2, host2 = make_engine(cname, n)
                variants[cname] = e2.kernel_variant()
                e2.upload_primaries_ptr(n, host2.data_ptr())
                e2.launch_resident(MASTER_SEED, rank * n)          # warm-up
                e2.synchronize()
                e2.reset_tallies()
                flush.zero_()
                torch.cuda.synchronize()
                e2.launch_resident(MASTER_SEED, (world + rank) * n)
                e2.synchronize()
                ms, st = e2.last_kernel_ms(), e2.counters()["steps"]
                pipelined = None
                if cname in PIPELINED_CONFIGS:
                    # The same launches alternating over TWO engines (own streams) of this GPU, as apps/mytrim_uo2 runs
                    # its chunks: the CTAs of the next launch take the SMs that the tail of the previous one leaves idle
                    # (work is shared inside a CTA only).  Host clock around asynchronous launches + synchronize.
                    e3, host3 = make_engine(cname, n)
                    e3.upload_primaries_ptr(n, host3.data_ptr())
                    e3.launch_resident(MASTER_SEED, rank * n)      # warm-up
                    e3.synchronize()
                    pair, n_l = (e2, e3), 4
                    for e in pair:
                        e.reset_tallies()
                        if cw["tally"] & capi.TALLY_IONLOG:
                            lib.mtb_clear_lists(e._h)
                    flush.zero_()
                    torch.cuda.synchronize()
                    tp = time.perf_counter()
                    for i in range(n_l):
                        e = pair[i % 2]
                        if i >= 2:
                            e.synchronize()
                            if cw["tally"] & capi.TALLY_IONLOG:
                                lib.mtb_clear_lists(e._h)
                        e.launch_resident(MASTER_SEED, ((2 + i) * world + rank) * n)
                    for e in pair:
                        e.synchronize()
                    pipelined = [(time.perf_counter() - tp) * 1e3 / n_l, float(sum(e.counters()["steps"] for e in pair)) / n_l]
                    e3.close()
                e2.close()
            if cname == name or cname not in PIPELINED_CONFIGS:
                pipelined = None
            v = torch.tensor([ms, float(st)] + (pipelined or [0.0, 0.0]), dtype=torch.float64, device="cuda")
            if world > 1:
                vmax, vsum = v.clone(), v.clone()
                dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(vsum, op=dist.ReduceOp.SUM)
            else:
                vmax, vsum = v, v
            t = float(vmax[0]) * 1e-3
            configs[cname] = {"cascades_per_s": n * world / t, "collision_steps_per_s": float(vsum[1]) / t,
                              "steps_per_cascade": float(vsum[1]) / (n * world), "primaries_per_gpu": n,
                              "kernel_ms": float(vmax[0]), "what": cw["desc"],
                              "kernel_variant": variants.get(cname, "") + (" (launches without records: MONO-NOREC)"
                                                                           if variants.get(cname) == "MONO" else "")}
            if pipelined:
                tp2 = float(vmax[2]) * 1e-3
                configs[cname]["two_engines_per_gpu"] = {
                    "cascades_per_s": n * world / tp2, "collision_steps_per_s": float(vsum[3]) / tp2, "ms_per_launch": float(vmax[2]),
                    "launches": 4, "clock": "host clock around asynchronous launches alternating over two engines + synchronize"}

    if world > 1:
        barrier()
    if rank != 0:
        if world > 1:
            # rank 0 may still exercise the C-ABI join on GPUs 0 and 1: wait on the host, not in an NCCL kernel
            dist.destroy_process_group()
        return 0

    # ---------------- roofline + CPU baseline (rank 0) ----------------
    tf = ctypes.c_double()
    ms = ctypes.c_float()
    lib.mtb_measure_fp32_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)]
    peak_src = "measured (fp32_peak_kernel, FFMA x 148 SMs, this run)"
    if lib.mtb_measure_fp32_peak(local_rank, ctypes.byref(tf), ctypes.byref(ms)) != 0 or tf.value <= 0:
        tf.value = FP32_LANES * 2 * 1.965e9 / 1e12
        peak_src = "nominal 148 SM x 128 lanes x 2 x 1965 MHz (probe failed)"
    steps_per_s_gpu0 = counters["steps"] / (kernel_ms * 1e-3)
    achieved = steps_per_s_gpu0 * FLOP_PER_STEP / 1e12
    # Counters that cannot be measured outside a profiler come from the committed ncu --set full capture of the same
    # kernel on the same workload (profiles/traffic.json, profiles/ncu_metrics.json, written by tools/ncu_traffic.py
    # and tools/ncu_metrics.py) and are scaled to this run's launch size and rate; the line says so.
    traffic, traffic_src = None, None
    tj = load_profile_json("traffic.json")
    if tj and name == HEADLINE:
        traffic = tj["dram_bytes_per_cascade"] * B
        traffic_src = "profiles/traffic.json (ncu --set full, %s), bytes per cascade x this launch size" % tj.get("source", "?")
    roofline = {
        "bound": "fp32-issue", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
        "traffic": traffic, "traffic_source": traffic_src, "kernel": "transport_kernel",
        "flop_per_collision_step": FLOP_PER_STEP,
        "collision_steps_per_launch": counters["steps"] / args.steps,
        "kernel_ms_per_launch": kernel_ms / args.steps, "peak_source": peak_src,
        "note": "no dense contraction on this path (SURVEY.md §8d): work = 930 FP32-equivalent flop per collision "
                "step (the reference's operation count); HBM traffic is O(100 B) per cascade",
    }
    mj = load_profile_json("ncu_metrics.json")
    if mj and name == HEADLINE and clocks.get("sm_mhz"):
        # machine-side view: thread instructions the kernel really executes per second against the FP32 lanes
        thread_inst_per_s = mj["thread_inst_per_cascade"] * (B * args.steps) / (kernel_ms * 1e-3)
        roofline.update({
            "issue_frac": thread_inst_per_s / (FP32_LANES * clocks["sm_mhz"] * 1e6),
            "thread_inst_per_collision_step": mj["thread_inst_per_cascade"] / (counters["steps"] / (B * args.steps)),
            "warp_efficiency": mj["warp_execution_efficiency"],
            "issue_slots_busy_pct": mj.get("issue_slots_busy_pct"),
            "issue_frac_source": "instruction counts from profiles/ncu_metrics.json (ncu --set full, %s) x this run's "
                                 "cascade rate / (148 x 128 lanes x SM clock under load)" % mj.get("source", "?")})
    line = {
        "metric": "cascades_per_s", "value": value, "unit": "cascades/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": job_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (+f64 position/energy accumulators)", "data": "synthetic",
        "config": {"workload": name, "description": wl["desc"], "primaries_per_gpu_per_step": B,
                   "l2": "256 MB buffer written between timed iterations", "tally_reduction_ms": red_ms},
        "collision_steps_per_s": coll_steps / (job_ms * 1e-3),
        "steps_per_cascade": coll_steps / cascades,
        "vacancies_per_ion": total["vacancies_created"] / cascades,
        "ions_per_cascade": total["ions"] / cascades,
        "recoils_per_s": (total["ions"] - total["primaries"]) / (job_ms * 1e-3),
        "wall_s_resident_arm": wall_resident,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "cascades/s", "h2d_bytes_per_step": int(B * capi.ION_DTYPE.itemsize),
                "d2h_bytes_per_step": int(d2h[0]), "includes_tally_join": world > 1,
                "primaries_in_joined_tallies": e2e_primaries_total},
        "gpu_launches": 2 * args.steps,
        "roofline": roofline,
    }
    if configs:
        line["configs"] = configs
    if world > 1:
        dist.destroy_process_group()  # the other ranks have left: GPUs 0 and 1 are free for the single-process check
        line["reduction_check"] = reduction_check
        line["c_abi_allreduce_check"] = c_abi_allreduce_check()
    if not args.no_cpu_baseline:
        from tests import util
        g.build_test_infrastructure()
        cores = os.cpu_count() or 1
        if name == "uo2_fission" and os.path.exists(util.REF_UO2):
            r = reference_uo2_primaries_per_s(1, cores)
            line["cpu_baseline"] = {
                "value": r["cascades_per_s"], "unit": "cascades/s", "cores": cores, "kind": "reference",
                "sample": "%d fission fragments, unmodified apps/mytrim_uo2.C, one process per core, %.1f s" % (
                    r["n"], r["seconds"])}
        elif name != "uo2_fission" and util.have_reference():
            n_ref = 5 * REF_PER_CORE[name] * cores   # a few seconds of wall time
            r = reference_cascades_per_s(name, n_ref, cores)
            line["cpu_baseline"] = {
                "value": r["cascades_per_s"], "unit": "cascades/s", "cores": cores, "kind": "reference",
                "sample": "%d cascades of %s, unmodified reference library on %d threads, %.1f s" % (
                    n_ref, name, cores, r["seconds"])}
        elif name != "uo2_fission":
            n_ref = 300
            orc = util.OracleEngine(util.ORC_RNG_PHILOX, tally_mask=capi.TALLY_VAC_DEPTH)
            c = util.setup_engine(orc, name)
            t0 = time.perf_counter()
            orc.run(util.primaries_for(c, n_ref), seed=MASTER_SEED)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_ref / dt, "unit": "cascades/s", "cores": 1, "kind": "port",
                                    "sample": "%d cascades, oracle C restatement, 1 thread" % n_ref}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
