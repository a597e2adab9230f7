#!/usr/bin/env python3
"""mtb_allreduce (the C-ABI tally join of a single-process multi-GPU job: dlopen'd NCCL, ncclCommInitAll) on GPUs 0
and 1: two handles, primaries sharded by global index, joined tallies equal one GPU running everything.  Prints
"ok ..." on success.  bench.py runs it in a child process with a time limit at N > 1."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from mytrim_b200 import capi, workloads  # noqa: E402

lib = capi.load_library()
if lib.mtb_device_count() < 2:
    print("skipped: one GPU")
    sys.exit(0)
cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH)
c = workloads.CONFIGS["cu_on_cu_10keV"]
n = 20000
ions = workloads.primaries_for(c, n)
with capi.Engine(device=0, **cfg) as one, capi.Engine(device=0, **cfg) as a, capi.Engine(device=1, **cfg) as b:
    for e in (one, a, b):
        workloads.setup_engine(e, c)
    one.run(ions, seed=5)
    a.run(ions[:n // 2], seed=5, first_index=0)
    b.run(ions[n // 2:], seed=5, first_index=n // 2)
    arr = (C.c_void_p * 2)(a._h, b._h)
    lib.mtb_allreduce.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    rc = lib.mtb_allreduce(arr, 2)
    if rc != 0:
        print("failed: mtb_allreduce status %d: %s" % (rc, lib.mtb_last_error().decode()))
        sys.exit(1)
    c1, ca, cb = one.counters(), a.counters(), b.counters()
    bad = [k for k in ("vacancies_created", "replacements", "steps", "ions", "primaries") if not c1[k] == ca[k] == cb[k]]
    if abs(c1["EelTotal"] - ca["EelTotal"]) > 1e-9 * c1["EelTotal"]:
        bad.append("EelTotal")
    if not (np.array_equal(one.vac_depth()[0], a.vac_depth()[0]) and np.array_equal(one.vac_depth()[1], b.vac_depth()[1])):
        bad.append("histograms")
    if bad:
        print("failed: " + ",".join(bad))
        sys.exit(1)
    print("ok: mtb_allreduce over 2 handles == 1 GPU (%d cascades, %d collision steps)" % (n, c1["steps"]))
