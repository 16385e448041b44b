#!/usr/bin/env python
"""Top stall-sample SASS instructions per kernel from an .ncu-rep (run where ncu is installed, no GPU needed).
usage: tools/ncu_hot.py report.ncu-rep [kernel-substring] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernel = None
blocks = {}
hdr = None
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kernel = row[1]
        blocks.setdefault(kernel, [])
        hdr = None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if hdr and kernel:
        blocks[kernel].append(dict(zip(hdr, row)))
seen = set()
for k, rows in blocks.items():
    if sub not in k or k in seen:
        continue
    seen.add(k)
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"== {k}  instructions={len(rows)} samples={tot}")
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        print(f"{i:5d} {int(r['# Samples'] or 0):5d} {100*int(r['# Samples'] or 0)/max(tot,1):5.1f}%  {r['Source'].strip()[:90]}")
