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
ap.add_argument("--escale", type=float, default=1.0, help="scale the primary energies (short kernels for ncu)")
args = ap.parse_args()



def uo2_fission_primaries(n, seed=39172):
    """Fission-fragment pairs like apps/mytrim_uo2 (mytrim_uo2.C:226-266): light/heavy mass peaks, ~170 MeV
    shared by momentum conservation, back-to-back isotropic directions, uniform origins."""
    import numpy as np
    rng = np.random.default_rng(seed)
    ne = (n + 1) // 2
    a1 = np.clip(rng.normal(96.0, 6.0, ne), 70.0, 117.0)
    a2 = 235.0 - a1
    etot = rng.normal(170.0e6, 8.0e6, ne)
    d = rng.normal(size=(ne, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    pos = rng.uniform(0.0, 400.0, (ne, 3))
    ions = capi.make_ions(2 * ne, 1, 1.0, 1.0)
    for k, (a, sgn) in enumerate(((a1, 1.0), (a2, -1.0))):
        ions["m"][k::2] = np.round(a, 3)
        ions["Z"][k::2] = np.rint(a * 92.0 / 235.0)
        ions["E"][k::2] = etot * (235.0 - a) / 235.0
        ions["pos"][k::2] = pos
        ions["dir"][k::2] = sgn * d
    return ions[:n]


with capi.Engine(tally_mask=args.tally) as eng:
    if args.workload == "uo2_fission":
        import numpy as np
        cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
        eng.set_materials([util.UO2, util.XE_GAS])
        eng.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
        ions = uo2_fission_primaries(args.primaries)
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
