#!/bin/bash
# persistent one-CTA-per-SM length-gauge step for ONE simulation (ensemble.cuh SOLO) vs one CTA per pair; parity from the bench's own check
tools/ab_env.sh c3_len 1000 "X=1" "ION_NO_SOLO=1" "X=1" "ION_NO_SOLO=1"
timeout 300 python -m pytest tests -m gpu -x -q -k "bench_shapes or len" 2>&1 | tail -3
