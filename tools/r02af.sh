#!/bin/bash
# C2 (LineMesh CN, 1024 x 2^16): segment width vs CTAs per SM
tools/ab_env.sh c2_line_ensemble 200 "X=1" "ION_TSEG=128" "ION_TSEG=192" "ION_TSEG=256" "ION_TSEG=320"
