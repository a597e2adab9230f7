#!/usr/bin/env python3
"""Aggregates tools/ncu_hotspots.py output by code region of the transport kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "TraitsT<(unsigned int)2048, (unsigned int)1>"  # north-star variant (MONO)
sass = sys.argv[3] if len(sys.argv) > 3 else "TraitsTILj2048ELj1E"
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hotspots.py"), rep, "--kernel", kernel, "--sass-kernel", sass,
                      "--top", "5000"],
                     capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(ROOT, "mytrim_b200", "csrc", "mtb_transport.cuh")).read().split("\n")
psrc = open(os.path.join(ROOT, "mytrim_b200", "csrc", "mtb_physics.cuh")).read().split("\n")


def find(lines, s):
    for i, l in enumerate(lines):
        if s in l:
            return i + 1
    return 0


tmarks = [("lookup", "lookup_cluster(const LaunchParams"), ("lane/class helpers", "struct Lane"), ("stack_store", "stack_store(StackEntry"),
          ("stack_load", "stack_load(const StackEntry"), ("log_birth", "log_birth(const LaunchParams"),
          ("finish_ion", "finish_ion(const LaunchParams"), ("depth_tally", "depth_tally(const"),
          ("vacancy_creation", "vacancy_creation(const"), ("close_subtree", "close_subtree(const"),
          ("work-sharing pool", "vload(const unsigned long long"), ("stack cursor / suspend_ion", "stack_entry(const LaunchParams"),
          ("block_add", "block_add(unsigned long long"), ("variant selection", "needed_features(const"), ("ionlog", "ionlog_append(const"),
          ("wrap_cell", "wrap_cell(int j"),
          ("loop: setup", "lane_loop(const LaunchParams"), ("loop: refill", "refill: next suspended"),
          ("loop: geometry+norm", "one collision: trim.C:74-424"), ("loop: philox+uniforms", "the four uniforms of this step"),
          ("loop: class/flight", "const float E0 = L.Ecur"), ("loop: element pick+pair", "target element — trim.C:147-156"),
          ("loop: stopping+scatter calls", "const float see = material_stopping"), ("loop: energy", "energy bookkeeping — trim.C"),
          ("loop: move+rotate", "recoil is born at the previous collision site"), ("loop: CUT", "CUT boundaries — trim.C:344-352"),
          ("loop: fate", "fate of recoil and projectile"), ("loop: events", "    if (EVENTS)"), ("loop: who flies next", "who flies next")]
pmarks = [("proton_stopping", "proton_stopping(const DevElement"), ("element_stopping", "element_stopping(const ProjClass"),
          ("material_stopping", "material_stopping(const ProjClass"), ("magic: newton step", "screening_sums(int potential"),
          ("magic: rutherford+guess", "magic_scatter(int potential"),
          ("magic: newton loop", "  do\n"), ("magic: tail", "// trim.C:235-271"), ("flight_from_pair", "flight_from_pair(const PairM"),
          ("on-the-fly pair", "make_pair_m(const ProjClass")]


def marks(lines, ms):
    r = [(n, find(lines, s)) for n, s in ms]
    r = sorted([m for m in r if m[1] > 0], key=lambda m: m[1])
    return r + [("end", len(lines) + 1)]


tm = marks(src, tmarks)
pm = [(n, find(psrc, s.strip("\n")) if s != "  do\n" else find(psrc, "  do")) for n, s in pmarks]
pm = sorted([m for m in pm if m[1] > 0], key=lambda m: m[1]) + [("end", len(psrc) + 1)]
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for l in out[2:]:
    m = re.match(r"\s*([\d.]+)%\s+([\d.]+)%\s+([\d.]+)\s+(\S+):(\d+)", l)
    if not m:
        continue
    s, i, t, f, ln = float(m.group(1)), float(m.group(2)), float(m.group(3)), m.group(4), int(m.group(5))
    key = f
    for fname, mk, tag in (("mtb_transport.cuh", tm, "T "), ("mtb_physics.cuh", pm, "P ")):
        if f == fname:
            key = tag + "?"
            for k in range(len(mk) - 1):
                if mk[k][1] <= ln < mk[k + 1][1]:
                    key = tag + mk[k][0]
                    break
    agg[key][0] += s
    agg[key][1] += i
    agg[key][2] += i * t
print(out[0])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s samples %5.1f%%  inst %5.1f%%  thr %4.1f" % (k, v[0], v[1], v[2] / max(v[1], 1e-9)))
