#!/usr/bin/env python
"""Executed warp-instructions and stall samples per CUDA SOURCE LINE of one kernel: joins the SASS page of an .ncu-rep
with the line table of the built library (nvdisasm --print-line-info; instruction order is identical).
usage: tools/ncu_lines.py report.ncu-rep mangled-kernel-substring demangled-substring [top-n] [lib.so]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, sub, dsub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ionization_b200", "_lib", "libionization_b200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
lines_of = []  # per instruction (in order) -> (file, line) of the innermost inlined location
cur, infn = None, False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        infn = sub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if "inlined at" not in ln or cur is None or True:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_of.append(cur)

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernel, hdr, rows = None, None, []
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        if rows:
            break
        kernel, hdr = row[1], None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if hdr and kernel and dsub in kernel:
        rows.append(dict(zip(hdr, row)))
if len(rows) != len(lines_of):
    print(f"warning: {len(rows)} profiled instructions vs {len(lines_of)} disassembled", file=sys.stderr)
inst, samp = collections.Counter(), collections.Counter()
for d, loc in zip(rows, lines_of):
    inst[loc] += int(d["Instructions Executed"] or 0)
    samp[loc] += int(d["# Samples"] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
src_cache = {}
def text(loc):
    if loc is None:
        return ""
    f, n = loc
    for d in ("ionization_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:100] if n - 1 < len(src_cache[p]) else ""
    return ""
print(f"kernel {kernel}: {ti} warp-instructions, {ts} samples")
for loc, n in inst.most_common(top):
    print(f"{100 * n / ti:5.1f}% inst {100 * samp[loc] / max(ts, 1):5.1f}% stall  {loc}  {text(loc)}")

if os.environ.get("BY_OP"):
    byop = collections.Counter()
    for d, loc in zip(rows, lines_of):
        src = d["Source"].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        byop[(loc, src.split()[0].split(".")[0])] += int(d["Instructions Executed"] or 0)
    print("---- (line, opcode)")
    for (loc, op), n in byop.most_common(int(os.environ["BY_OP"])):
        print(f"{100 * n / ti:5.1f}%  {op:8s} {loc}  {text(loc)[:80]}")
