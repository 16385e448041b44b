#!/bin/bash
python -m pytest tests -m gpu -x -q -k "eight_rows or fused_observation or datastores or observ" 2>&1 | tail -3
python tools/obs_probe.py VEL 512
python tools/obs_probe.py VEL 512
