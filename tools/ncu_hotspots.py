#!/usr/bin/env python3
"""Per-source-line hot spots of a kernel from an ncu report (needs -lineinfo + --set full).

    python tools/ncu_hotspots.py gpurun_out/r01_prof.ncu-rep [--kernel transport_kernel] [--top 40]

Joins `ncu --page source --csv` (per-SASS-instruction samples and executed counts) with the line
table `nvdisasm -g` prints for the cubin inside mytrim_b200/libmytrim_b200.so (the library must be
the build that was profiled)."""
import argparse
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--kernel", default="transport_kernel")
ap.add_argument("--sass-kernel", default="", help="substring of the mangled name in the cubin (default: --kernel)")
ap.add_argument("--lib", default=os.path.join(ROOT, "mytrim_b200", "libmytrim_b200.so"))
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--by", default="line", choices=["line", "file", "op"])
args = ap.parse_args()

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", args.lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

addr2line, addr2op = {}, {}
inside, cur = False, ("?", 0)
for l in dis.split("\n"):
    if l.startswith("\t.section\t.text."):
        inside = (args.sass_kernel or args.kernel) in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        a = int(m.group(1), 16)
        addr2line[a] = cur
        op = m.group(2).split()
        op = op[1] if op[0].startswith("@") else op[0]
        addr2op[a] = op.split(".")[0]

raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
agg = defaultdict(lambda: [0, 0, 0])
base = None
for r in rows:
    if r and r[0] == "Kernel Name":
        hdr = None
        active = args.kernel in r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        base = None
        continue
    if hdr is None or not active or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        addr = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
    except ValueError:
        continue
    if base is None:
        base = addr
    off = addr - base
    samples = int(d["# Samples"] or 0)
    inst = int(d["Instructions Executed"] or 0)
    tinst = int(d["Thread Instructions Executed"] or 0)
    if args.by == "op":
        key = addr2op.get(off, "?")
    else:
        f, ln = addr2line.get(off, ("?", 0))
        key = f if args.by == "file" else "%s:%d" % (f, ln)
    a = agg[key]
    a[0] += samples
    a[1] += inst
    a[2] += tinst

ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("kernel %s: %d samples, %d warp instructions, %.2f avg active threads" % (
    args.kernel, ts, ti, sum(a[2] for a in agg.values()) / ti))
print("%8s %8s %6s  %s" % ("samples%", "inst%", "thr", "where"))
srccache = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
    text = ""
    if args.by == "line" and ":" in key:
        f, ln = key.rsplit(":", 1)
        p = os.path.join(ROOT, "mytrim_b200", "csrc", f)
        if os.path.exists(p):
            if p not in srccache:
                srccache[p] = open(p).read().split("\n")
            if 0 < int(ln) <= len(srccache[p]):
                text = srccache[p][int(ln) - 1].strip()[:80]
    print("%7.2f%% %7.2f%% %6.1f  %-26s %s" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / max(a[1], 1), key, text))
