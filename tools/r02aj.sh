#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
tools/ab_env.sh c2_line_ensemble 300 "X=1" "ION_LINE_M=4"
