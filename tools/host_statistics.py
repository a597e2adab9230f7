#!/usr/bin/env python3
"""North-star statistical criterion WITHOUT a GPU: 1e6 Cu->Cu 10 keV cascades through the host build of the device
loop (tests/libhostsim.so, same FP32 algorithm and Philox streams as the kernels; all host cores, ~1.5 min on 8)
against the 1e6-cascade summary of the unmodified reference (tests/golden/ref_stats_cu_on_cu_10keV_1e6.npz).

    python tools/host_statistics.py [--n 1000000] > profiles/r01_statistics_host_loop_1e6.log
"""
import argparse
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mytrim_b200 import capi  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--seed", type=int, default=2344)
args = ap.parse_args()
W = os.cpu_count() or 1
PER = (args.n + W - 1) // W


def work(i):
    c = util.CONFIGS["cu_on_cu_10keV"]
    n = min(PER, args.n - i * PER)
    with util.HostSimEngine(tally_mask=capi.TALLY_RECORDS) as hs:
        util.setup_engine(hs, c)
        return hs.run(util.primaries_for(c, n), seed=args.seed, first_index=i * PER, records=True)


if __name__ == "__main__":
    t0 = time.time()
    with Pool(W) as pool:
        rec = np.concatenate(pool.map(work, [i for i in range(W) if i * PER < args.n]))
    print("host build of the device loop: %d Cu->Cu 10 keV cascades on %d cores in %.0f s (seed %d, global indices 0..n-1: "
          "the same streams the GPU suite's 1e6-ion test uses)" % (len(rec), W, time.time() - t0, args.seed))
    summary = np.load(os.path.join(util.GOLDEN, "ref_stats_cu_on_cu_10keV_1e6.npz"))
    print("reference: %d cascades of the unmodified library, distinct 32-bit seeds" % int(summary["n"]))
    ok = True
    for k, (a, b, D, p) in util.ks_against_summary(rec, summary).items():
        good = p > 0.01 and abs(a - b) <= 0.01 * abs(b)
        ok &= good
        print("%-24s mean %12.5f vs %12.5f (%+.4f %%)  KS D=%.5f p=%.3f  %s" % (k, a, b, 100 * (a / b - 1), D, p, "ok" if good else "FAIL"))
    print("PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
