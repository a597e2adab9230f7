#!/usr/bin/env python3
"""Extracts dram__bytes_read.sum + dram__bytes_write.sum of the transport kernel from an ncu --set full
report and writes profiles/traffic.json (bytes per cascade), which bench.py scales to its launch size.

    python tools/ncu_traffic.py gpurun_out/r01_prof.ncu-rep <cascades in the profiled launch>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, n = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if "transport_kernel" not in d.get("Kernel Name", ""):
        continue
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k]) * scale[units[hdr.index(k)]]
    out = {"kernel": d["Kernel Name"], "cascades_in_profiled_launch": n, "dram_bytes": tot,
           "dram_bytes_per_cascade": tot / n, "source": os.path.basename(rep),
           "duration_ms": float(d["gpu__time_duration.sum"]) * (1e-6 if units[hdr.index("gpu__time_duration.sum")] == "ns" else 1.0)}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)
    break
