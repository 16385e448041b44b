#!/bin/bash
# k_slab code-shape variants (pass loop unrolled; mask loaded late) on the headline
D=/root/repo/ionization_b200/_lib
tools/ab_env.sh c3_vel 1000 "X=1" "ION_LIB=$D/exp_up.so" "ION_LIB=$D/exp_lm.so" "ION_LIB=$D/exp_uplm.so"
