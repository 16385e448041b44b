#!/bin/bash
# C2 (LineMesh CN), eight rows per thread: register cap 128 (two 256-thread CTAs per SM) vs 156 registers
A=/root/repo/ionization_b200/_lib/exp_m8.so
B=/root/repo/ionization_b200/_lib/exp_m8b.so
tools/ab_env.sh c2_line_ensemble 200 "ION_LIB=$B ION_M=8 ION_TSEG=192" "ION_LIB=$B ION_M=8 ION_TSEG=128" "ION_LIB=$B ION_M=8 ION_TSEG=96" "ION_LIB=$A ION_M=8 ION_TSEG=64" "ION_LIB=$A ION_M=8 ION_TSEG=128"
