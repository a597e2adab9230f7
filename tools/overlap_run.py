#!/usr/bin/env python3
"""Launches of a work-sharing workload back to back on ONE engine against the same launches alternating over TWO
engines (own streams) of the same GPU: the CTAs of the next launch fill the SMs that the tail of the previous one
leaves idle.  Wall clock around asynchronous launches + synchronize (the kernels overlap, their own timers do not add).

    python tools/overlap_run.py --workload uo2_fission --primaries 65536 --launches 4
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mytrim_b200 import capi, workloads  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--primaries", type=int, default=1 << 16)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--workload", default="uo2_fission")
ap.add_argument("--tally", type=int, default=capi.TALLY_IONLOG)
ap.add_argument("--engines", type=int, default=2)
args = ap.parse_args()


def make():
    eng = capi.Engine(tally_mask=args.tally, ionlog_z=54)
    if args.workload == "uo2_fission":
        ions = workloads.setup_workload(eng, "uo2_fission", args.primaries)
    else:
        c = util.CONFIGS[args.workload]
        util.setup_engine(eng, c)
        ions = util.primaries_for(c, args.primaries)
    eng.upload_primaries(ions)
    return eng


engs = [make() for _ in range(args.engines)]
for e in engs:  # warm-up
    e.launch_resident(2344, 0)
for e in engs:
    e.synchronize()
    e.reset_tallies()

# (a) one engine, back to back
t = time.perf_counter()
for i in range(args.launches):
    engs[0].launch_resident(2344, (i + 1) * args.primaries)
    engs[0].synchronize()
seq = time.perf_counter() - t
steps_seq = engs[0].counters()["steps"]
for e in engs:
    e.reset_tallies()

# (b) the same launches alternating over the engines: launch i+E is queued as soon as launch i is done
t = time.perf_counter()
pending = [False] * len(engs)
for i in range(args.launches):
    k = i % len(engs)
    if pending[k]:
        engs[k].synchronize()
    engs[k].launch_resident(2344, (i + 1) * args.primaries)
    pending[k] = True
for e in engs:
    e.synchronize()
ovl = time.perf_counter() - t
steps_ovl = sum(e.counters()["steps"] for e in engs)
n = args.launches * args.primaries
print("%s n=%d x %d launches: one engine %.1f ms/launch (%.3e steps/s), %d engines %.1f ms/launch (%.3e steps/s), %+.1f %%; steps equal: %s" % (
    args.workload, args.primaries, args.launches, 1e3 * seq / args.launches, steps_seq / seq, len(engs), 1e3 * ovl / args.launches,
    steps_ovl / ovl, 100.0 * (ovl / seq - 1.0), steps_seq == steps_ovl))
for e in engs:
    e.close()
