#!/usr/bin/env python3
"""Statistical criterion of the north star: projected range, lateral range, vacancies per ion and the
energy partition of the CUDA path against the UNMODIFIED reference on the same inputs — two-sample KS
p > 0.01 and means within 1 % at N ions (default 1e6).

    python tools/validate_statistics.py [--n 1000000] [--workload cu_on_cu_10keV] [--out profiles/...json]

The reference runs through oracle/_ref/ref_driver on all host cores with DISTINCT 32-bit per-primary
seeds (stock runmytrim draws 16-bit seeds, so its samples repeat beyond 65536 primaries)."""
import argparse
import json
import os
import sys
import time

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mytrim_b200 import capi  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--workload", default="cu_on_cu_10keV")
ap.add_argument("--out", default="")
args = ap.parse_args()

c = util.CONFIGS[args.workload]
n = args.n
cores = os.cpu_count() or 1
t0 = time.time()
ref, summary, _ = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"], util.distinct_seeds(n),
                                              tally="phonon", threads=cores, box=c.get("box"), timeout=2400)
t_ref = time.time() - t0

runs = {}
for label, mask in (("lean variant with energy partition (PHONON|RECORDS)", capi.TALLY_PHONON | capi.TALLY_RECORDS),
                    ("fast kernel (VAC_DEPTH|RECORDS)", capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)):
    with capi.Engine(tally_mask=mask) as eng:
        util.setup_engine(eng, c)
        ions = util.primaries_for(c, n)
        t0 = time.time()
        rec = eng.run(ions, seed=20261017, records=True)
        runs[label] = (rec, time.time() - t0, eng.last_kernel_ms())

box = c.get("box")
cy, cz = (box[1] / 2, box[2] / 2) if box else (50.0, 50.0)
fields = {
    "projected_range_x": lambda r: r["pos"][:, 0],
    "lateral_range": lambda r: np.hypot(r["pos"][:, 1] - cy, r["pos"][:, 2] - cz),
    "vacancies_per_ion": lambda r: r["vacancies"].astype(float),
    "replacements_per_ion": lambda r: r["replacements"].astype(float),
    "electronic_loss_Eel": lambda r: r["Eel"],
    "nuclear_loss_Enuc": lambda r: r["Enuc"],
    "collision_steps": lambda r: r["steps"].astype(float),
    "ions_followed": lambda r: r["ions"].astype(float),
}
report = {"workload": args.workload, "n": n, "reference": {"seconds": t_ref, "threads": cores, "summary": summary},
          "criterion": "two-sample KS p > 0.01 and |mean_gpu/mean_ref - 1| < 0.01", "runs": {}}
ok = True
for label, (rec, wall, kms) in runs.items():
    out = {"wall_s": wall, "kernel_ms": kms, "fields": {}}
    for name, get in fields.items():
        if name == "nuclear_loss_Enuc" and "fast" in label:
            continue
        a, b = get(rec), get(ref)
        ks = stats.ks_2samp(a, b)
        rel = float(a.mean() / b.mean() - 1.0)
        passed = bool(ks.pvalue > 0.01 and abs(rel) < 0.01)
        ok &= passed
        out["fields"][name] = {"mean_gpu": float(a.mean()), "mean_ref": float(b.mean()), "rel_mean_diff": rel,
                               "ks_D": float(ks.statistic), "ks_p": float(ks.pvalue), "pass": passed}
        print("%-34s %-24s mean %12.5g vs %12.5g (%+.4f%%)  KS D=%.5f p=%.3f  %s" % (
            label, name, a.mean(), b.mean(), 100 * rel, ks.statistic, ks.pvalue, "ok" if passed else "FAIL"))
    report["runs"][label] = out
report["pass"] = bool(ok)
if args.out:
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
print("PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
