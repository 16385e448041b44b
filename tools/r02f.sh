#!/bin/bash
o=gpurun_out
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_slab -s 60 -c 1 -f -o $o/r02f_slab_obs python tools/obs_probe.py VEL 24 > $o/r02f_a.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_unit -s 62 -c 1 -f -o $o/r02f_len_obs python tools/obs_probe.py LEN 24 > $o/r02f_b.log 2>&1
ls -la $o/r02f*
