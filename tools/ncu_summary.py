#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) and a --set full report into markdown.
usage: tools/ncu_summary.py launches.csv report.ncu-rep > profiles/rNN_ncu_summary.md"""
import csv
import collections
import subprocess
import sys


def short(name):
    name = name.replace("ion::", "").replace("(int)", "")
    progs = {"0": "ROT", "1": "ROT_CN_ROT", "2": "H2", "3": "H2_CN_H2", "4": "CN", "5": "LINE_SO_LEN", "6": "LINE_SO_VEL", "7": "LINE_CN", "8": "LEN_STEP"}
    if name.startswith("void k_unit<") or name.startswith("k_unit<"):
        args = name.split("<")[1].split(">")[0].split(",")
        return f"k_unit<M={args[0].strip()}, {progs.get(args[1].strip(), args[1].strip())}, TMAX={args[2].strip()}>"
    return name.split("(")[0].replace("void ", "")


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = rows[hdr]
    kn, mv, mu = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1 :]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        if r[mu] in ("ns", "nsecond"):
            v /= 1e3
        elif r[mu] in ("ms", "msecond"):
            v *= 1e3
        k = short(r[kn])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def full_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = [
        ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__waves_per_multiprocessor", "waves/SM"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("lts__t_bytes.sum", "L2 bytes"),
    ]
    idx = [(hdr.index(m), label) for m, label in want if m in hdr]
    res = []
    for r in rows[2:]:
        d = collections.OrderedDict(kernel=short(r[hdr.index("Kernel Name")]))
        for i, label in idx:
            d[label] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    return res


def main():
    launches, report = sys.argv[1], sys.argv[2]
    agg = launch_list(launches)
    total = sum(v[1] for v in agg.values())
    print("## Launch list (ncu `--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total µs | mean µs | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / total:.1f} % |")
    print("\n## `ncu --set full` capture (one launch of each kernel of a time step)\n")
    recs = full_report(report)
    seen = set()
    keys = list(recs[0].keys())
    print("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys))
    for d in recs:
        if d["kernel"] in seen:
            continue
        seen.add(d["kernel"])
        print("| " + " | ".join(f"`{v}`" if k == "kernel" else v for k, v in d.items()) + " |")


if __name__ == "__main__":
    main()
