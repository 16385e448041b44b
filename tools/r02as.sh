#!/bin/bash
# C4 ensemble: psi prefetch by the TMA engine (cp.async.bulk + mbarrier, ION_ENS_BULK=1) vs 16 cp.async per thread
tools/ab_env.sh c4_len_ensemble 300 "X=1" "ION_ENS_BULK=1" "X=1" "ION_ENS_BULK=1"
ION_ENS_BULK=1 timeout 300 python -m pytest tests -m gpu -x -q -k "ensemble or bench_shapes" 2>&1 | tail -2
