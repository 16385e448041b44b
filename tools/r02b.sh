#!/bin/bash
L=$PWD/ionization_b200/_lib
tools/ab_env.sh c3_vel 1000 "X=1" "ION_SLAB_NT=288" "ION_SLAB_G=16" "ION_SLAB_G=16 ION_SLAB_NT=288" "ION_LIB=$L/exp_pair64.so" "ION_LIB=$L/exp_pair64.so ION_SLAB_NT=288"
tools/ab_env.sh c3_len 1000 "X=1" "ION_NO_LINEAR_ANGLES=1" "ION_LIB=$L/exp_pair64.so"
tools/ab_env.sh c4_len_ensemble 200 "X=1" "ION_NO_LINEAR_ANGLES=1"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
