#!/usr/bin/env python3
"""bench.py — full ion cascades per second on the north-star workload.

Workload (BASELINE.json / SURVEY.md §8d config 1): Cu (Z=29, m=63.546) ions at 10 keV into a
1000 A Cu layer (rho 8.92), full recoil cascades (follow ALL), TrimVacCount tallies.  A "step" is
one batch of `--primaries` cascades per GPU through the transport kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

value   = cascades/s with the primaries already resident in HBM (device time of the K launches,
          CUDA events on the launching stream, max over ranks, plus the per-step tally all-reduce
          when N > 1)
e2e     = the same metric through mtb_run() with HOST buffers: every step copies its primaries
          host->device (pinned memory) and reads the tallies back.
The reference arm (--impl reference) times the UNMODIFIED reference library (oracle/_ref) on all
host cores on a bounded sample of the same workload.
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
WORKLOAD = "cu_on_cu_10keV"
WORKLOAD_DESC = "Cu->Cu 10 keV, 1000 A Cu layer, full cascades, TrimVacCount tallies (validation/cu_on_cu)"
MASTER_SEED = 2344


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


def reference_cascades_per_s(n, threads, timeout=900):
    """Times oracle/_ref/ref_driver (the unmodified reference library, runmytrim-equivalent set-up)."""
    from tests import util
    c = util.CONFIGS[WORKLOAD]
    lines = util.reference_script(c["ion"], c["materials"], c["thicknesses"], n=n, tally="vaccount",
                                  threads=threads, master=MASTER_SEED)
    lines.append("run")
    out = util.run_reference("\n".join(lines) + "\n", timeout=timeout)
    return json.loads(out[-1])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build_test_infrastructure()
    from tests import util
    cores = os.cpu_count() or 1
    if not util.have_reference():
        emit({"impl": "reference", "unavailable": "oracle/_ref was not built (reference tree absent)"})
        return 0
    per_step = args.ref_cascades if args.ref_cascades else 400 * cores
    for _ in range(args.warmup):
        reference_cascades_per_s(max(per_step // 8, cores), cores)
    t_total, n_total, steps_total = 0.0, 0, 0
    for _ in range(args.steps):
        r = reference_cascades_per_s(per_step, cores)
        t_total += r["seconds"]
        n_total += r["n"]
        steps_total += r["steps"]
    value = n_total / t_total
    sample = "%d steps x %d Cu->Cu 10 keV cascades on %d threads (unmodified reference, TrimVacCount)" % (
        args.steps, per_step, cores)
    line = {
        "impl": "reference", "metric": "cascades_per_s", "value": value, "unit": "cascades/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "description": WORKLOAD_DESC, "primaries_per_step": per_step},
        "collision_steps_per_s": steps_total / t_total,
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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--primaries", type=int, default=1 << 23, help="cascades per GPU per step")
    ap.add_argument("--ref-cascades", type=int, default=0, help="cascades per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

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
    from tests import util

    B = args.primaries
    c = util.CONFIGS[WORKLOAD]
    eng = capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH, device=local_rank)
    util.setup_engine(eng, c)

    # primaries in pinned host memory (the e2e arm copies them every step)
    pinned = torch.empty(B * capi.ION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    ions = np.frombuffer(pinned.numpy(), dtype=capi.ION_DTYPE)
    ions[:] = util.primaries_for(c, B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # tallies as torch tensors over the engine's device memory (for the NCCL reduction)
    from mytrim_b200 import dist as mdist
    t_u64, t_f64 = mdist.tally_tensors(eng)

    def reduce_tallies():
        """Only the additive tallies cross NVLink (mytrim_b200/dist.py): the analogue of threadJoin."""
        if world == 1:
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mdist.reduce_tallies(t_u64, t_f64)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def first_index(step):
        return (step * world + rank) * B

    # ---------------- resident arm: `value` ----------------
    eng.upload_primaries_ptr(B, pinned.data_ptr())
    eng.synchronize()
    step_id = 0
    for _ in range(args.warmup):
        eng.launch_resident(MASTER_SEED, first_index(step_id))
        eng.synchronize()
        step_id += 1
    if world > 1:
        # warm-up of the collective itself (NCCL builds its communicator lazily on first use)
        for _ in range(args.warmup):
            mdist.reduce_tallies(torch.zeros_like(t_u64), torch.zeros_like(t_f64))
    eng.reset_tallies()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, red_ms = 0.0, 0.0
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the CUDA-event brackets)
        torch.cuda.synchronize()
        eng.launch_resident(MASTER_SEED, first_index(step_id))
        eng.synchronize()
        dev_ms += eng.last_kernel_ms()
        step_id += 1
    barrier()
    wall_resident = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    counters = eng.counters()  # this rank's tallies over the timed steps
    red_ms = reduce_tallies()  # whole-job tallies (one reduction per job, as in runmytrim's threadJoin)
    total = eng.counters() if world > 1 else counters
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
    nb = ctypes.c_size_t()
    cnt = capi.Counters()
    lib = capi.load_library()

    def e2e_step(step):
        rc = lib.mtb_run(eng._h, B, pinned.data_ptr(), MASTER_SEED, first_index(step), None)
        if rc != 0:
            raise RuntimeError(lib.mtb_last_error().decode())
        lib.mtb_get_counters(eng._h, ctypes.byref(cnt))
        lib.mtb_get_vac_depth(eng._h, vac_host.ctypes.data, repl_host.ctypes.data, len(vac_host), ctypes.byref(nb))

    e2e_step(step_id)
    step_id += 1
    eng.reset_tallies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step(step_id)
        step_id += 1
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = cascades / float(t_e[0])
    d2h = ctypes.sizeof(capi.Counters) + vac_host.nbytes + repl_host.nbytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline + CPU baseline (rank 0) ----------------
    tf = ctypes.c_double()
    ms = ctypes.c_float()
    lib.mtb_measure_fp32_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)]
    peak_src = "measured (fp32_peak_kernel, FFMA x 148 SMs, this run)"
    if lib.mtb_measure_fp32_peak(local_rank, ctypes.byref(tf), ctypes.byref(ms)) != 0 or tf.value <= 0:
        tf.value = 148 * 128 * 2 * 1.965e9 / 1e12
        peak_src = "nominal 148 SM x 128 lanes x 2 x 1965 MHz (probe failed)"
    steps_per_s_gpu0 = counters["steps"] / (kernel_ms * 1e-3)
    achieved = steps_per_s_gpu0 * FLOP_PER_STEP / 1e12
    # DRAM traffic of the kernel: bytes per cascade from the committed ncu --set full capture
    # (profiles/traffic.json, written by tools/ncu_traffic.py), scaled to this launch size
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_cascade"] * B
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "bound": "fp32-issue", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
        "traffic": traffic, "kernel": "transport_kernel", "flop_per_collision_step": FLOP_PER_STEP,
        "collision_steps_per_launch": counters["steps"] / args.steps,
        "kernel_ms_per_launch": kernel_ms / args.steps, "peak_source": peak_src,
        "note": "no dense contraction on this path (SURVEY.md §8d): work = 930 FP32-equivalent flop per collision "
                "step; HBM traffic is O(100 B) per cascade",
    }
    line = {
        "metric": "cascades_per_s", "value": value, "unit": "cascades/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": job_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (+f64 position/energy accumulators)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "description": WORKLOAD_DESC, "primaries_per_gpu_per_step": B,
                   "l2": "256 MB buffer written between timed iterations", "tally_reduction_ms": red_ms},
        "collision_steps_per_s": coll_steps / (job_ms * 1e-3),
        "steps_per_cascade": coll_steps / cascades,
        "vacancies_per_ion": total["vacancies_created"] / cascades,
        "ions_per_cascade": total["ions"] / cascades,
        "recoils_per_s": (total["ions"] - total["primaries"]) / (job_ms * 1e-3),
        "wall_s_resident_arm": wall_resident,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "cascades/s", "h2d_bytes_per_step": int(B * capi.ION_DTYPE.itemsize),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": 2 * args.steps,
        "roofline": roofline,
    }
    if not args.no_cpu_baseline:
        g.build_test_infrastructure()
        cores = os.cpu_count() or 1
        if util.have_reference():
            n_ref = 2000 * cores   # ~2 s of wall time, ~30 core-seconds
            r = reference_cascades_per_s(n_ref, cores)
            line["cpu_baseline"] = {
                "value": r["cascades_per_s"], "unit": "cascades/s", "cores": cores, "kind": "reference",
                "sample": "%d Cu->Cu 10 keV cascades, unmodified reference library on %d threads, %.1f s" % (
                    n_ref, cores, r["seconds"])}
        else:
            n_ref = 300
            orc = util.OracleEngine(util.ORC_RNG_PHILOX, tally_mask=capi.TALLY_VAC_DEPTH)
            util.setup_engine(orc, c)
            t0 = time.perf_counter()
            orc.run(util.primaries_for(c, n_ref), seed=MASTER_SEED)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_ref / dt, "unit": "cascades/s", "cores": 1, "kind": "port",
                                    "sample": "%d cascades, oracle C restatement, 1 thread" % n_ref}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
