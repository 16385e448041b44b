#!/bin/bash
python -m pytest tests -m gpu -x -q -k "ensemble or config4" 2>&1 | tail -3
tools/ab_env.sh c4_len_ensemble 200 "X=1" "X=2"
