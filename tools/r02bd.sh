#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/obs_probe.py VEL 512
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench c3_vel us/step', d['us_per_time_step'], 'observed', d['extra']['observed_every_step'])"
