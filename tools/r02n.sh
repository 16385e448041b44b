#!/bin/bash
o=gpurun_out
python -m pytest tests -m gpu -x -q -k "fused or observed or datastores or config3" 2>&1 | tail -3
python tools/obs_probe.py VEL 512; python tools/obs_probe.py LEN 512
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_unit -s 62 -c 1 -f -o $o/r02n_len_obs python tools/obs_probe.py LEN 24 > $o/r02n_b.log 2>&1
python tools/ncu_lines.py $o/r02n_len_obs.ncu-rep k_unitILi4ELi9ELi512ELb0 "k_unit<(int)4, (int)9, (int)512, (bool)0>" 45 > $o/r02n_lines_len_obs.txt 2>&1
python tools/ncu_stalls.py $o/r02n_len_obs.ncu-rep > $o/r02n_len_obs.stalls.txt 2>&1
rm -f $o/r02n_len_obs.ncu-rep
cat $o/r02n_len_obs.stalls.txt | head -16; head -48 $o/r02n_lines_len_obs.txt
