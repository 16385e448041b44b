#!/usr/bin/env python
"""Per-kernel stall-reason breakdown (pc sampling) and key utilisation metrics from an .ncu-rep.
usage: tools/ncu_stalls.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
stall = [(i, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum", "dram__bytes_read.sum", "l1tex__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__waves_per_multiprocessor", "sm__ctas_launched.sum"]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if name in seen:
        continue
    seen.add(name)
    print("==", name[:70])
    for k in keys:
        if k in hdr:
            print(f"   {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
    tot = sum(float(r[i] or 0) for i, _ in stall)
    parts = sorted(((float(r[i] or 0), n) for i, n in stall), reverse=True)
    print("   stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in parts[:8] if v > 0))
