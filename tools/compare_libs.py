#!/usr/bin/env python3
"""Runs the same primaries through two builds of the library (tools/build_variants.sh) and compares per-primary
records, depth histograms and counters bit for bit: a scheduling change (hand-over batching, launch bounds, ...)
must not change any result.   python tools/compare_libs.py build/variants/base.so build/variants/x.so [workload n]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) >= 2 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    from mytrim_b200 import capi, workloads
    name, n, out = sys.argv[2], int(sys.argv[3]), sys.argv[4]
    with capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS) as eng:
        c = workloads.setup_engine(eng, name)
        rec = eng.run(workloads.primaries_for(c, n), seed=2344, records=True)
        vac, repl = eng.vac_depth()
        cnt = eng.counters()
    np.savez(out, rec=rec, vac=vac, repl=repl, cnt=np.array([cnt[k] for k in sorted(cnt) if k != "stack_max"], dtype=np.float64))
    sys.exit(0)

import numpy as np  # noqa: E402
a, b = sys.argv[1], sys.argv[2]
name = sys.argv[3] if len(sys.argv) > 3 else "cu_on_cu_10keV"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
res = []
for k, lib in enumerate((a, b)):
    out = "/tmp/compare_libs_%d.npz" % k
    subprocess.run([sys.executable, __file__, "--child", name, str(n), out], check=True,
                   env=dict(os.environ, MYTRIM_B200_LIB=os.path.abspath(lib)))
    res.append(np.load(out))
ra, rb = res[0]["rec"], res[1]["rec"]
bad = [f for f in ra.dtype.names if not np.array_equal(ra[f], rb[f])]
same_h = np.array_equal(res[0]["vac"], res[1]["vac"]) and np.array_equal(res[0]["repl"], res[1]["repl"])
same_c = np.allclose(res[0]["cnt"], res[1]["cnt"], rtol=1e-13, atol=0)
print("%s vs %s on %d x %s: records %s, histograms %s, counters %s" % (
    os.path.basename(a), os.path.basename(b), n, name, "identical" if not bad else "DIFFER in " + ",".join(bad),
    "identical" if same_h else "DIFFER", "equal" if same_c else "DIFFER"))
sys.exit(0 if (not bad and same_h and same_c) else 1)
