#!/bin/bash
# ADI: closing radial solve with eight rows per thread (k_adi_r) vs k_unit<PROG_CN>
python -m pytest tests -m gpu -x -q -k "adi or ADI" 2>&1 | tail -3
tools/ab_env.sh c3_adi 1000 "X=1" "ION_NO_ADI_R=1" "X=1"
