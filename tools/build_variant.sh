#!/bin/bash
# build an A/B variant of the engine: tools/build_variant.sh NAME [-DFLAG ...]  ->  ionization_b200/_lib/exp_NAME.so  (use with ION_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ionization_b200/_lib
nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -o ionization_b200/_lib/exp_$name.so ionization_b200/csrc/engine.cu
echo ionization_b200/_lib/exp_$name.so
