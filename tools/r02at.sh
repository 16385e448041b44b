#!/bin/bash
# LU factors of the pair kernels staged by the TMA engine (two cp.async.bulk per CTA + mbarrier) vs 8 cp.async per thread (exp_nolub.so)
N=/root/repo/ionization_b200/_lib/exp_nolub.so
tools/ab_env.sh c3_vel 1000 "X=1" "ION_LIB=$N" "X=1" "ION_LIB=$N"
tools/ab_env.sh c3_len 1000 "X=1" "ION_LIB=$N"
tools/ab_env.sh c1_len 2000 "X=1" "ION_LIB=$N"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
