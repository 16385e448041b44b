#!/bin/bash
# C4 ensemble: members per item (LU factors staged once per item)
D=/root/repo/ionization_b200/_lib
tools/ab_env.sh c4_len_ensemble 300 "X=1" "ION_LIB=$D/exp_mb32.so" "ION_LIB=$D/exp_mb64.so" "ION_LIB=$D/exp_mb8.so" "X=1"
