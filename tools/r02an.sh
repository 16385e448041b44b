#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/obs_probe.py LEN 512
python tools/obs_probe.py VEL 512
