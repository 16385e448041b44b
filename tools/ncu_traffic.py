#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in an `ncu --set full` report -> profiles/ncu_traffic.json (read by bench.py's
roofline.traffic).  usage: tools/ncu_traffic.py WORKLOAD report.ncu-rep [profiles/ncu_traffic.json]"""
import csv
import re
import json
import os
import subprocess
import sys

workload, rep = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
PROGS = {"0": "rot", "1": "rot_cn_rot", "2": "h2", "3": "h2_cn_h2", "4": "cn", "5": "line_so_len", "6": "line_so_vel", "7": "line_cn", "8": "len_step", "9": "len_step_obs", "10": "len_step_halo"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("ion::", "").replace("(int)", "").replace("(bool)", "").replace("void ", "")
    if name.startswith("k_unit<"):
        return f"k_unit<{PROGS.get(name.split('<')[1].split(',')[1].strip(), '?')}>"
    return name.split("<")[0].split("(")[0]


rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
kn, rd, wr, du = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
# FP64 instructions executed (thread level): add `--metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,
# smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum` to the capture
fp = [i for i, h in enumerate(hdr) if re.search(r"sass_thread_inst_executed_op_d(fma|mul|add)_pred_on\.sum$", h)]
acc = {}
for r in rows[2:]:
    k = short(r[kn])
    a = acc.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[4] += sum(float(r[i].replace(",", "")) for i in fp)
    a[1] += float(r[rd].replace(",", "")) * UNIT[units[rd]]
    a[2] += float(r[wr].replace(",", "")) * UNIT[units[wr]]
    a[3] += float(r[du].replace(",", ""))
try:
    data = json.load(open(out_path))
except Exception:
    data = {}
data[workload] = {k: {"dram_bytes_read": a[1] / a[0], "dram_bytes_write": a[2] / a[0], "launches_captured": a[0], **({"fp64_thread_inst": a[4] / a[0]} if fp else {}), "source": os.path.basename(rep) + " (ncu --set full, caches flushed between replays)"} for k, a in acc.items()}
json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(data[workload], indent=1))
