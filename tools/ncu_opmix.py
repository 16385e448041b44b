#!/usr/bin/env python
"""Dynamic instruction mix per kernel (warp-level executed counts per SASS opcode) from an .ncu-rep source page.
usage: tools/ncu_opmix.py report.ncu-rep [kernel-substring]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernel, hdr, done = None, None, set()
agg = collections.defaultdict(lambda: collections.Counter())
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kernel = row[1] if row[1] not in done else None
        hdr = None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if row[0] in ("File Path", "Function Name", "Line No"):
        if kernel:
            done.add(kernel)
        kernel = None
        continue
    if hdr and kernel:
        d = dict(zip(hdr, row))
        op = d["Source"].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        agg[kernel][op.split(".")[0]] += int(d["Instructions Executed"] or 0)
for k, c in agg.items():
    if sub not in k:
        continue
    tot = sum(c.values())
    print(f"== {k}  warp-instructions executed: {tot}")
    for op, n in c.most_common(22):
        print(f"   {op:12s} {n:10d} {100 * n / tot:5.1f}%")
