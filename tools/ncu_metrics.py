#!/usr/bin/env python3
"""From an ncu --set full report of the transport kernel: (1) the raw counter page as CSV into profiles/ (the
reproducible source of the markdown summary), (2) profiles/ncu_metrics.json — warp and thread instruction counts per
cascade, warp execution efficiency, issue-slot utilisation — which bench.py turns into roofline.issue_frac.

    python tools/ncu_metrics.py gpurun_out/r02_prof.ncu-rep <cascades in the profiled launch> [tag]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, n = sys.argv[1], int(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
with open(os.path.join(ROOT, "profiles", "%s_ncu_raw.csv" % tag), "w") as f:
    f.write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if "transport_kernel" not in d.get("Kernel Name", ""):
        continue
    warp_inst = float(d["smsp__inst_executed.sum"])
    eff = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"])
    out = {"kernel": d["Kernel Name"], "cascades_in_profiled_launch": n, "source": os.path.basename(rep),
           "warp_inst": warp_inst, "warp_execution_efficiency": eff,
           "warp_inst_per_cascade": warp_inst / n, "thread_inst_per_cascade": warp_inst * eff / n,
           "issue_slots_busy_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
           "registers_per_thread": int(float(d["launch__registers_per_thread"])),
           "achieved_occupancy_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"])}
    with open(os.path.join(ROOT, "profiles", "ncu_metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)
    break
