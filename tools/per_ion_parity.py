"""Prints the per-ion deterministic-criterion counts (tests/util.py::compare_ion_logs) of every case of
tests/parity_cases.py: CUDA kernels against the FP32 host replay and against the FP64 oracle.  Needs a B200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mytrim_b200 import capi  # noqa: E402
from tests import parity_cases, util  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
tot = {}
for name, n in parity_cases.PER_ION_CASES:
    n = max(1, int(n * scale))
    cfg = dict(tally_mask=capi.TALLY_IONLOG | capi.TALLY_RECORDS, ionlog_capacity=1 << 23, **parity_cases.case_options(name))
    with capi.Engine(**cfg) as eng, util.HostSimEngine(**cfg) as hs, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        ions = parity_cases.setup_case(eng, name, n)
        parity_cases.setup_case(hs, name, n)
        parity_cases.setup_case(orc, name, n)
        rg = eng.run(ions, seed=2344, records=True)
        rh = hs.run(ions, seed=2344, records=True)
        ro = orc.run(ions, seed=2344, records=True)
        lg = eng.ion_log(1 << 23)
        for partner, log, rec in (("fp32_replay", hs.ion_log(1 << 23), rh), ("fp64_oracle", orc.ion_log(1 << 23), ro)):
            s = util.compare_ion_logs(lg, log, ions)
            r = util.compare_records(rg, rec, ions)
            print("%-18s n=%-5d vs %-11s ions %s | cascades %s" % (name, n, partner, s, r), flush=True)
            t = tot.setdefault(partner, dict(ions=0, joined=0, ints_equal=0, pos_outliers=0, energy_outliers=0, own_energy_outliers=0))
            t["ions"] += max(s["n_test"], s["n_replay"])
            for k in ("joined", "ints_equal", "pos_outliers", "energy_outliers", "own_energy_outliers"):
                t[k] += s[k]
for partner, t in tot.items():
    print("TOTAL vs %s: %s  identical share %.6f  position outlier share %.2e" %
          (partner, t, t["ints_equal"] / t["ions"], t["pos_outliers"] / max(t["joined"], 1)))
