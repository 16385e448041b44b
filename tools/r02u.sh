#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tools/ab_env.sh c3_adi 1000 "X=1" "ION_ADI_PW=2" "ION_ADI_PW=8"
tools/ab_env.sh c3_vel 1000 "X=1"
tools/ab_env.sh c3_len 1000 "X=1"
