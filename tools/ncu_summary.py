#!/usr/bin/env python3
"""Condenses an ncu --set full report into the handful of counters the north star names
(FP32/SFU pipe utilisation, warp execution efficiency, L1/shared hit rate, achieved occupancy)
plus issue/stall and DRAM traffic numbers.   python tools/ncu_summary.py report.ncu-rep [kernel-substr]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else "transport_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid size (CTAs)"),
    ("launch__block_size", "block size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps/SM)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (active threads / 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (FP32) utilisation (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe utilisation (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU/SFU pipe utilisation (%)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe utilisation (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe utilisation (%)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX hit rate (%)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("sass__inst_executed_local_loads", "local-memory loads"),
    ("sass__inst_executed_local_stores", "local-memory stores"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "uniform branch targets (%)"),
]
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if sub not in d.get("Kernel Name", ""):
        continue
    print("## %s" % d["Kernel Name"])
    print()
    print("| counter | value | unit |")
    print("|---|---|---|")
    for k, label in KEYS:
        if k in d:
            print("| %s (`%s`) | %s | %s |" % (label, k, d[k], units[hdr.index(k)]))
    print()
    print("Stall reasons (average warps stalled per issue-active cycle):")
    print()
    st = sorted(((float(d[h]), h) for h in STALLS if d.get(h)), reverse=True)
    for v, h in st[:10]:
        print("- %s: %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    print()
