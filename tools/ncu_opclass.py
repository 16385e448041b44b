#!/usr/bin/env python
"""Executed warp-instructions by SASS opcode (and stall samples) for one kernel of an .ncu-rep captured with --import-source on.
usage: tools/ncu_opclass.py report.ncu-rep [kernel-substring]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernel, hdr, done = None, None, set()
ops = collections.Counter()
samp = collections.Counter()
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        if kernel and sub in kernel:
            done.add(kernel)
        kernel, hdr = row[1], None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if hdr and kernel and sub in kernel and kernel not in done:
        d = dict(zip(hdr, row))
        src = d["Source"].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0]
        full = src.split()[0]
        if op in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST", "SHFL"):
            op = full if op in ("LDS", "STS", "LDL", "STL") and False else op
        ops[op] += int(d["Instructions Executed"] or 0)
        samp[op] += int(d["# Samples"] or 0)
tot, stot = sum(ops.values()), sum(samp.values())
print(f"total warp-instructions {tot}, samples {stot}")
for op, n in ops.most_common(40):
    print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(stot, 1):5.1f}%")
