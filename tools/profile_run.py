#!/usr/bin/env python3
"""Small driver for ncu: a few resident launches of the north-star workload, nothing else.

    ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
        -o gpurun_out/prof python tools/profile_run.py --primaries 262144 --launches 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mytrim_b200 import capi  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--primaries", type=int, default=1 << 18)
ap.add_argument("--launches", type=int, default=2)
ap.add_argument("--workload", default="cu_on_cu_10keV")
ap.add_argument("--tally", type=int, default=capi.TALLY_VAC_DEPTH)
ap.add_argument("--ionlog-z", type=int, default=54, help="Z filter of the ion log (tally bit 64)")
ap.add_argument("--escale", type=float, default=1.0, help="scale the primary energies (short kernels for ncu)")
args = ap.parse_args()



with capi.Engine(tally_mask=args.tally, ionlog_z=args.ionlog_z) as eng:
    if args.workload == "uo2_fission":
        from mytrim_b200 import workloads
        ions = workloads.setup_workload(eng, "uo2_fission", args.primaries)   # mtb_fission_pairs: the app's source
        ions["E"] *= args.escale
        eng.upload_primaries(ions)
    else:
        c = util.CONFIGS[args.workload]
        util.setup_engine(eng, c)
        ions = util.primaries_for(c, args.primaries)
        ions["E"] *= args.escale
        eng.upload_primaries(ions)
    for i in range(args.launches):
        eng.launch_resident(2344, i * args.primaries)
        eng.synchronize()
        cnt = eng.counters()
        ms = eng.last_kernel_ms()
        print("launch %d: %.3f ms, %.3e cascades/s, %.3e steps/s (cumulative steps %d)" % (
            i, ms, args.primaries / (ms * 1e-3), cnt["steps"] / (i + 1) / (ms * 1e-3), cnt["steps"]))
