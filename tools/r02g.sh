#!/bin/bash
ION_LEN4=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
tools/ab_env.sh c3_len 1000 "X=1" "ION_LEN4=1"
tools/ab_env.sh c4_len_ensemble 200 "X=1" "ION_LEN4=1"
tools/ab_env.sh c1_len 2000 "X=1" "ION_LEN4=1"
