#!/usr/bin/env python3
"""North-star statistical criterion WITHOUT a GPU: cascades through the host build of the device loop
(tests/libhostsim.so, same FP32 algorithm and Philox streams as the kernels; all host cores, 1e6 Cu->Cu 10 keV
cascades take ~1.5 min on 8) against the committed summary of the unmodified reference's cascades for that
configuration (tests/golden/ref_stats_<workload>.npz, tests/golden/make_golden.py::STATISTICS_CASES).

    python tools/host_statistics.py [--workload cu_on_cu_10keV] [--n <as many as the reference sample>]
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
ap.add_argument("--workload", default="cu_on_cu_10keV")
ap.add_argument("--n", type=int, default=0, help="cascades (default: the size of the reference sample)")
ap.add_argument("--seed", type=int, default=2344)
args = ap.parse_args()
SUMMARY = np.load(os.path.join(util.GOLDEN, "ref_stats_%s.npz" % args.workload))
if args.n <= 0:
    args.n = int(SUMMARY["n"])
W = os.cpu_count() or 1
PER = (args.n + W - 1) // W


def work(i):
    c = util.CONFIGS[args.workload]
    n = min(PER, args.n - i * PER)
    with util.HostSimEngine(tally_mask=capi.TALLY_RECORDS) as hs:
        util.setup_engine(hs, c)
        return hs.run(util.primaries_for(c, n), seed=args.seed, first_index=i * PER, records=True)


if __name__ == "__main__":
    t0 = time.time()
    with Pool(W) as pool:
        rec = np.concatenate(pool.map(work, [i for i in range(W) if i * PER < args.n]))
    print("host build of the device loop: %d %s cascades on %d cores in %.0f s (seed %d, global indices 0..n-1: "
          "the same streams the GPU suite's statistical test uses)" % (len(rec), args.workload, W, time.time() - t0, args.seed))
    summary = SUMMARY
    print("reference: %d cascades of the unmodified library, distinct 32-bit seeds" % int(summary["n"]))
    ok = True
    for k, (a, b, D, p) in util.ks_against_summary(rec, summary).items():
        good = p > 0.01 and abs(a - b) <= 0.01 * abs(b)
        ok &= good
        print("%-24s mean %12.5f vs %12.5f (%+.4f %%)  KS D=%.5f p=%.3f  %s" % (k, a, b, 100 * (a / b - 1), D, p, "ok" if good else "FAIL"))
    print("PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
