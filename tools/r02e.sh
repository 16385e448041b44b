#!/bin/bash
o=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02e_obs_vel.csv python tools/obs_probe.py VEL 24 > $o/r02e_obs_vel.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02e_obs_len.csv python tools/obs_probe.py LEN 24 > $o/r02e_obs_len.log 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/r02e_obs_vel.csv", "gpurun_out/r02e_obs_len.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        a = acc.setdefault(r[kn][:60], [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
    print(f)
    for k, (c, t) in acc.items(): print(f"  {k:60s} n={c:4d} mean {t / c / 1e3:8.2f} us")
PY
