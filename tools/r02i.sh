#!/bin/bash
tools/ab_env.sh c3_vel 1000 "ION_NO_LATE=1" "X=1" "ION_NO_LATE=1" "X=1"
tools/ab_env.sh c3_len 1000 "ION_NO_LATE=1" "X=1" "ION_NO_LATE=1" "X=1"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
