#!/bin/bash
# C2 (LineMesh CN): eight rows per thread in r-segments of 256-thread CTAs (experiment build) vs the product
L=/root/repo/ionization_b200/_lib/exp_m8.so
tools/ab_env.sh c2_line_ensemble 200 "X=1" "ION_LIB=$L ION_M=8 ION_TSEG=192" "ION_LIB=$L ION_M=8 ION_TSEG=128" "ION_LIB=$L ION_M=8 ION_TSEG=160"
ION_LIB=$L ION_M=8 ION_TSEG=192 python bench.py --workload c2_line_ensemble --steps 1 --warmup 1 --time-steps 50 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-400
